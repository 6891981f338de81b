// walk_f32.cu -- fp32 CRW training walk (forward + fused reverse pass) for sm_100a.
//
// Replaces the tail of CRW.forward, /root/reference/src/model.py:22-46, and autograd through it
// (scripts/train.py:71):
//   model.py:22   F.normalize                      -> inverse row norms, folded into the GEMM epilogue
//   model.py:26   einsum(...)/tau                  -> walk_affinity_kernel (A_t, S_t, S'_t in one pass)
//   model.py:31   cat/flip/transpose palindrome    -> never materialised (index arithmetic)
//   model.py:35-44 (T-2)^2 softmax+bmm loop        -> O(T) chain  L_k = L_{k-1} S'_{k-1},  R_k = S_{k-1} R_{k-1},
//                                                    M_k = L_k R_k   
//   model.py:45   cross_entropy(M^T, I)            -> walk_cycle_kernel epilogue (lse - diag), also emits
//                                                    G_k = rowsoftmax(M_k) - I for the reverse pass
//   autograd                                       -> walk_bwd_* kernels (Appendix A.3)
//
// All N x N state lives in the `saved` workspace (L2-resident at the reference's sizes); every
// stage is built from one CTA-wide tiled fp32 GEMM (64x64 tile, 4x4 per thread).
#include <cstdlib>
#include "common.cuh"
#include "walk_layout.cuh"

namespace crw {

constexpr int kWT = 256;   // threads per CTA
constexpr int kBM = 64, kBN = 64, kBK = 16;
constexpr int kLdT = kBM + 4;

struct GemmSmem {
    float As[kBK][kLdT];
    float Bs[kBK][kLdT];
};

// C[m,n] = sum_k opA(A)[m,k] * opB(B)[k,n] (* kscale[k]).   !TA: A[m*lda+k], TA: A[k*lda+m];
// !TB: B[k*ldb+n], TB: B[n*ldb+k].  epi(m, n, value) is called once per valid output element.
template <bool TA, bool TB, class Epi>
__device__ __forceinline__ void cta_gemm(const float* A, int lda, const float* B, int ldb, int M, int Nn, int K,
                                         const float* kscale, GemmSmem& sm, Epi epi) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int m0 = 0; m0 < M; m0 += kBM)
        for (int n0 = 0; n0 < Nn; n0 += kBN) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
            for (int k0 = 0; k0 < K; k0 += kBK) {
                __syncthreads();
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int idx = tid + e * kWT;
                    {
                        int kk, r;
                        if (TA) { r = idx % kBM; kk = idx / kBM; } else { kk = idx % kBK; r = idx / kBK; }
                        const int gm = m0 + r, gk = k0 + kk;
                        float v = 0.0f;
                        if (gm < M && gk < K) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
                        sm.As[kk][r] = v;
                    }
                    {
                        int kk, c;
                        if (TB) { kk = idx % kBK; c = idx / kBK; } else { c = idx % kBN; kk = idx / kBN; }
                        const int gn = n0 + c, gk = k0 + kk;
                        float v = 0.0f;
                        if (gn < Nn && gk < K) {
                            v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
                            if (kscale) v *= kscale[gk];
                        }
                        sm.Bs[kk][c] = v;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < kBK; ++kk) {
                    const float4 a = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
                    const float4 b = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
                    if (m < M && n < Nn) epi(m, n, acc[i][j]);
                }
        }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward stage 1: grid (T-1, B).  inverse norms, A_t, S_t = rowsoftmax(A_t), S'_t = rowsoftmax(A_t^T)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWT) walk_affinity_kernel(const float* __restrict__ x, float* ws, float* A_out,
                                                            int B, int T, int N, int C, float inv_tau) {
    __shared__ GemmSmem sm;
    extern __shared__ float dyn[];   // inv norms of frame t and t+1: [2][N]
    const WalkLayout lay(B, T, N, C);
    const int t = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* x0 = x + ((size_t)b * T + t) * N * C;
    const float* x1 = x0 + (size_t)N * C;
    float* inv0 = dyn;
    float* inv1 = dyn + N;
    for (int r = warp; r < 2 * N; r += kWT / 32) {
        const float* xr = (r < N) ? x0 + (size_t)r * C : x1 + (size_t)(r - N) * C;
        float ss = 0.0f;
        for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
        if (lane == 0) {
            dyn[r] = inv;
            if (r < N) ws[lay.invn + ((size_t)b * T + t) * N + r] = inv;
            else if (t == T - 2) ws[lay.invn + ((size_t)b * T + t + 1) * N + (r - N)] = inv;
        }
    }
    __syncthreads();
    float* At = ws + lay.mat(lay.A, b, t);
    float* Ao = A_out ? A_out + ((size_t)b * (T - 1) + t) * N * N : nullptr;
    cta_gemm<false, true>(x0, C, x1, C, N, N, C, nullptr, sm, [&](int i, int j, float v) {
        const float a = v * inv0[i] * inv1[j] * inv_tau;
        At[(size_t)i * N + j] = a;
        if (Ao) Ao[(size_t)i * N + j] = a;
    });
    float* S = ws + lay.mat(lay.S, b, t);
    float* Sp = ws + lay.mat(lay.Sp, b, t);
    // rows of S_t (softmax over j) and rows of S'_t (softmax over i of A[i][j])
    for (int r = warp; r < 2 * N; r += kWT / 32) {
        const bool col = r >= N;
        const int i = col ? r - N : r;
        const size_t step = col ? (size_t)N : 1, base = col ? (size_t)i : (size_t)i * N;
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) mx = fmaxf(mx, At[base + j * step]);
        mx = warp_max(mx);
        float se = 0.0f;
        for (int j = lane; j < N; j += 32) se += __expf(At[base + j * step] - mx);
        se = warp_sum(se);
        const float inv = 1.0f / se;
        float* dst = (col ? Sp : S) + (size_t)i * N;
        for (int j = lane; j < N; j += 32) dst[j] = __expf(At[base + j * step] - mx) * inv;
    }
}

// ------------------------------------------------------------------------------------------
// forward stage 2: grid (2, B).  x = 0: L chain, x = 1: R chain (sequential in k)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWT) walk_chain_kernel(float* ws, int B, int T, int N, int C) {
    __shared__ GemmSmem sm;
    const WalkLayout lay(B, T, N, C);
    const int b = blockIdx.y, K = T - 2;
    const size_t nn = (size_t)N * N;
    if (blockIdx.x == 0) {
        float* L0 = ws + lay.mat(lay.L, b, 0);
        for (size_t i = threadIdx.x; i < nn; i += kWT) L0[i] = (i / N == i % N) ? 1.0f : 0.0f;
        __syncthreads();
        for (int k = 1; k <= K; ++k) {
            const float* Lp = ws + lay.mat(lay.L, b, k - 1);
            const float* Sp = ws + lay.mat(lay.Sp, b, k - 1);
            float* Lk = ws + lay.mat(lay.L, b, k);
            cta_gemm<false, false>(Lp, N, Sp, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { Lk[(size_t)i * N + j] = v; });
        }
    } else {
        float* R1 = ws + lay.mat(lay.R, b, 1);
        for (size_t i = threadIdx.x; i < nn; i += kWT) R1[i] = (i / N == i % N) ? 1.0f : 0.0f;
        __syncthreads();
        for (int k = 2; k <= K; ++k) {
            const float* Rp = ws + lay.mat(lay.R, b, k - 1);
            const float* S = ws + lay.mat(lay.S, b, k - 1);
            float* Rk = ws + lay.mat(lay.R, b, k);
            cta_gemm<false, false>(S, N, Rp, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { Rk[(size_t)i * N + j] = v; });
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward stage 3: grid (T-2, B).  M_k = L_k R_k, loss partial, G_k = rowsoftmax(M_k) - I
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWT) walk_cycle_kernel(float* ws, int B, int T, int N, int C) {
    __shared__ GemmSmem sm;
    __shared__ float red[kWT / 32];
    const WalkLayout lay(B, T, N, C);
    const int k = blockIdx.x + 1, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* Lk = ws + lay.mat(lay.L, b, k);
    const float* Rk = ws + lay.mat(lay.R, b, k);
    float* Gk = ws + lay.mat(lay.G, b, k);
    cta_gemm<false, false>(Lk, N, Rk, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { Gk[(size_t)i * N + j] = v; });
    float part = 0.0f;
    for (int d = warp; d < N; d += kWT / 32) {
        float* row = Gk + (size_t)d * N;
        float mx = -INFINITY;
        for (int c = lane; c < N; c += 32) mx = fmaxf(mx, row[c]);
        mx = warp_max(mx);
        float se = 0.0f;
        for (int c = lane; c < N; c += 32) se += __expf(row[c] - mx);
        se = warp_sum(se);
        const float diag = row[d];
        __syncwarp();
        const float inv = 1.0f / se;
        for (int c = lane; c < N; c += 32) row[c] = __expf(row[c] - mx) * inv - (c == d ? 1.0f : 0.0f);
        part += (logf(se) + mx) - diag;   // identical on all lanes
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < kWT / 32; ++w) s += red[w];
        ws[lay.part + (size_t)b * (T - 1) + k] = s;
    }
}

// forward stage 4: one warp, fixed summation order.  loss = sum_{b,k} part / (B*N) / N
__global__ void walk_loss_reduce_kernel(const float* ws, float* loss, int B, int T, int N, int C) {
    const WalkLayout lay(B, T, N, C);
    const int lane = threadIdx.x;
    float s = 0.0f;
    for (int i = lane; i < B * (T - 2); i += 32) {
        const int b = i / (T - 2), k = i % (T - 2) + 1;
        s += ws[lay.part + (size_t)b * (T - 1) + k];
    }
    s = warp_sum(s);
    if (lane == 0) *loss = s / ((float)B * (float)N) / (float)N;
}

__global__ void walk_zero_loss_kernel(float* loss) { *loss = 0.0f; }

// bwd stage 1: grid (T-2, B, 2).  z=0: dL_k = s * G_k R_k^T ;  z=1: dR_k = s * L_k^T G_k
__global__ void __launch_bounds__(kWT) walk_bwd_own_kernel(const float* ws, float* sc, const float* dloss, int B, int T,
                                                           int N, int C) {
    __shared__ GemmSmem sm;
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int k = blockIdx.x + 1, b = blockIdx.y;
    const float s = *dloss / ((float)B * (float)N * (float)N);
    const float* Gk = ws + lay.mat(lay.G, b, k);
    if (blockIdx.z == 0) {
        const float* Rk = ws + lay.mat(lay.R, b, k);
        float* o = sc + lay.mat(bl.dL, b, k);
        cta_gemm<false, true>(Gk, N, Rk, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { o[(size_t)i * N + j] = v * s; });
    } else {
        const float* Lk = ws + lay.mat(lay.L, b, k);
        float* o = sc + lay.mat(bl.dR, b, k);
        cta_gemm<true, false>(Lk, N, Gk, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { o[(size_t)i * N + j] = v * s; });
    }
}

// bwd stage 2: grid (2, B).  x=0: dL_j += dL_{j+1} S'_j^T (j = K-1..1);  x=1: dR_j += S_j^T dR_{j+1} (j = K-1..2)
__global__ void __launch_bounds__(kWT) walk_bwd_chain_kernel(const float* ws, float* sc, int B, int T, int N, int C) {
    __shared__ GemmSmem sm;
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int b = blockIdx.y, K = T - 2;
    if (blockIdx.x == 0) {
        for (int j = K - 1; j >= 1; --j) {
            const float* dLn = sc + lay.mat(bl.dL, b, j + 1);
            const float* Sp = ws + lay.mat(lay.Sp, b, j);
            float* o = sc + lay.mat(bl.dL, b, j);
            cta_gemm<false, true>(dLn, N, Sp, N, N, N, N, nullptr, sm, [&](int i, int c, float v) { o[(size_t)i * N + c] += v; });
        }
    } else {
        for (int j = K - 1; j >= 2; --j) {
            const float* dRn = sc + lay.mat(bl.dR, b, j + 1);
            const float* S = ws + lay.mat(lay.S, b, j);
            float* o = sc + lay.mat(bl.dR, b, j);
            cta_gemm<true, false>(S, N, dRn, N, N, N, N, nullptr, sm, [&](int i, int c, float v) { o[(size_t)i * N + c] += v; });
        }
    }
}

// bwd stage 3: grid (T-1, B).  dS'_t = L_t^T dL_{t+1}; dS_t = dR_{t+1} R_t^T; softmax backward; dA_t
__global__ void __launch_bounds__(kWT) walk_bwd_dA_kernel(const float* ws, float* sc, const float* dA_ext, int B, int T,
                                                          int N, int C) {
    __shared__ GemmSmem sm;
    extern __shared__ float dyn[];   // rS[N], rSp[N]
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int t = blockIdx.x, b = blockIdx.y, K = T - 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hasSp = (t + 1 <= K), hasS = (t >= 1 && t + 1 <= K);
    float* dSp = sc + lay.mat(bl.dSp, b, t);
    float* dS = sc + lay.mat(bl.dS, b, t);
    float* dA = sc + lay.mat(bl.dAw, b, t);
    const float* S = ws + lay.mat(lay.S, b, t);
    const float* Sp = ws + lay.mat(lay.Sp, b, t);
    const float* ext = dA_ext ? dA_ext + ((size_t)b * (T - 1) + t) * N * N : nullptr;
    if (hasSp) {
        const float* Lt = ws + lay.mat(lay.L, b, t);
        const float* dLn = sc + lay.mat(bl.dL, b, t + 1);
        cta_gemm<true, false>(Lt, N, dLn, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { dSp[(size_t)i * N + j] = v; });
    }
    if (hasS) {
        const float* Rt = ws + lay.mat(lay.R, b, t);
        const float* dRn = sc + lay.mat(bl.dR, b, t + 1);
        cta_gemm<false, true>(dRn, N, Rt, N, N, N, N, nullptr, sm, [&](int i, int j, float v) { dS[(size_t)i * N + j] = v; });
    }
    float* rS = dyn;
    float* rSp = dyn + N;
    for (int r = warp; r < 2 * N; r += kWT / 32) {
        const bool second = r >= N;
        const int i = second ? r - N : r;
        float acc = 0.0f;
        if (second ? hasSp : hasS) {
            const float* P = (second ? Sp : S) + (size_t)i * N;
            const float* dP = (second ? dSp : dS) + (size_t)i * N;
            for (int j = lane; j < N; j += 32) acc = fmaf(P[j], dP[j], acc);
            acc = warp_sum(acc);
        }
        if (lane == 0) dyn[r] = acc;
    }
    __syncthreads();
    for (size_t e = threadIdx.x; e < (size_t)N * N; e += kWT) {
        const int i = (int)(e / N), j = (int)(e % N);
        float g = ext ? ext[e] : 0.0f;
        if (hasS) g += S[e] * (dS[e] - rS[i]);
        if (hasSp) g += Sp[(size_t)j * N + i] * (dSp[(size_t)j * N + i] - rSp[j]);
        dA[e] = g;
    }
}

// bwd stage 4: grid (T, B).  dE_t = (dA_t E_{t+1} + dA_{t-1}^T E_{t-1}) / tau, then normalise backward
__global__ void __launch_bounds__(kWT) walk_bwd_dx_kernel(const float* __restrict__ x, const float* ws, const float* sc,
                                                          float* dx, int B, int T, int N, int C, float inv_tau) {
    __shared__ GemmSmem sm;
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int t = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* o = dx + ((size_t)b * T + t) * N * C;
    const float* invn = ws + lay.invn + (size_t)b * T * N;
    if (t <= T - 2) {
        const float* dA = sc + lay.mat(bl.dAw, b, t);
        const float* xn = x + ((size_t)b * T + t + 1) * N * C;
        cta_gemm<false, false>(dA, N, xn, C, N, C, N, invn + (size_t)(t + 1) * N, sm,
                               [&](int i, int c, float v) { o[(size_t)i * C + c] = v * inv_tau; });
    } else {
        for (size_t e = threadIdx.x; e < (size_t)N * C; e += kWT) o[e] = 0.0f;
        __syncthreads();
    }
    if (t >= 1) {
        const float* dAp = sc + lay.mat(bl.dAw, b, t - 1);
        const float* xp = x + ((size_t)b * T + t - 1) * N * C;
        cta_gemm<true, false>(dAp, N, xp, C, N, C, N, invn + (size_t)(t - 1) * N, sm,
                              [&](int i, int c, float v) { o[(size_t)i * C + c] += v * inv_tau; });
    }
    const float* xt = x + ((size_t)b * T + t) * N * C;
    for (int i = warp; i < N; i += kWT / 32) {
        const float inv = invn[(size_t)t * N + i];
        float* orow = o + (size_t)i * C;
        const float* xr = xt + (size_t)i * C;
        if (inv >= 1.0f / kNormEps) {   // ||x|| <= eps: F.normalize divides by the constant eps
            for (int c = lane; c < C; c += 32) orow[c] *= inv;
            continue;
        }
        float dot = 0.0f;
        for (int c = lane; c < C; c += 32) dot = fmaf(xr[c] * inv, orow[c], dot);
        dot = warp_sum(dot);
        for (int c = lane; c < C; c += 32) orow[c] = (orow[c] - xr[c] * inv * dot) * inv;
    }
}

}  // namespace crw

namespace crw {
bool walk_small_supported(int N, int C);
bool walk_small_mma_supported(int N, int C);
int walk_small_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st, bool mma);
bool walk_fused_supported(int N, int C, int T);
bool walk_fused_roles_apply(int B);
size_t walk_fused_saved_bytes(int B, int T);
int walk_fused_backward(const float* x, const void* saved, const float* dloss, const float* dA_or_null, int B, int T, int N, int C, float tau,
                        float* dx, void* scratch, cudaStream_t st);
size_t walk_fused_bwd_scratch_bytes(int B, int T);
int walk_fused_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, void* saved, cudaStream_t st);
int walk_small_backward(const float* x, const float* ws, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st, bool mma);
// walk_tc_tiles.cu: tile-parallel tcgen05 bf16x3 path
int walk_tiles_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st);
int walk_tiles_backward(const float* x, const float* ws, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st);
size_t walk_tiles_saved_extra_bytes(int B, int T, int N, int C);     // bf16 operand planes kept next to the fp32 state
size_t walk_tiles_scratch_extra_bytes(int B, int T, int N, int C);
}  // namespace crw

using namespace crw;

// BF16X3 at the reference's sizes (N <= 64, C = 128): the fused tcgen05 kernels (walk_fused.cu), one launch per direction.
//   4 B <= #SMs   role-split kernels: four CTAs per batch element (affinity producers / chain / cycle; chain / dA / per-frame dE) that
//                 hand operand tiles over through L2 -- measured at B = 32, T = 10, N = 47 (CUDA-graph replay): 0.097 ms against
//                 0.114 ms for the eight shared-memory kernels (0.132 ms fp32)
//   B >= 96       one CTA per batch element: its time does not grow with the batch until it exceeds the SM count (B = 128: 0.220
//                 vs 0.284 ms)
//   in between    the eight shared-memory kernels with warp-level MMAs (walk_small.cu), which spread every stage over B x T CTAs
// CRW_WALK_FUSED=1 / 0 forces the choice.  (Forward and backward see the same B, T, N, C and environment, hence the same layout
// of `saved`.)
static bool walk_use_fused(int B, int T, int N, int C) {
    if (!walk_fused_supported(N, C, T) || getenv("CRW_WALK_FORCE_TILES")) return false;      // the test switch for the tile engine wins
    const char* e = getenv("CRW_WALK_FUSED");
    if (e) return atoi(e) != 0;
    return B >= 96 || walk_fused_roles_apply(B);
}

static inline bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" size_t crw_walk_saved_bytes(int B, int T, int N, int C, int precision) {
    if (B < 1 || T < 2 || N < 1 || C < 1) return 0;
    size_t b = WalkLayout(B, T, N, C).total * sizeof(float) + 256;
    if (precision == CRW_PREC_BF16X3) b += walk_tiles_saved_extra_bytes(B, T, N, C) + 256;
    if (precision == CRW_PREC_BF16X3 && walk_fused_supported(N, C, T)) { const size_t f = walk_fused_saved_bytes(B, T); if (f > b) b = f; }
    return b;
}

extern "C" size_t crw_walk_backward_scratch_bytes(int B, int T, int N, int C, int precision) {
    if (B < 1 || T < 2 || N < 1) return 0;
    size_t b = BwdLayout(B, T, N).total * sizeof(float) + 256;
    if (precision == CRW_PREC_BF16X3) b += walk_tiles_scratch_extra_bytes(B, T, N, C) + 256;
    if (precision == CRW_PREC_BF16X3 && walk_fused_supported(N, C, T)) { const size_t f = walk_fused_bwd_scratch_bytes(B, T); if (f > b) b = f; }
    return b;
}

static inline float* align256(void* p) {
    return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
}

extern "C" int crw_walk_forward(const float* x, int B, int T, int N, int C, float tau, int precision, float* loss,
                                float* A_or_null, void* saved, size_t saved_bytes, void* stream) {
    if (!x || !loss || B < 1 || T < 2 || N < 1 || C < 1 || !(tau > 0.0f)) return CRW_ERR_INVALID;
    if (precision != CRW_PREC_FP32 && precision != CRW_PREC_BF16X3) return CRW_ERR_UNSUPPORTED;
    if (!saved || saved_bytes < crw_walk_saved_bytes(B, T, N, C, precision)) return CRW_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = align256(saved);
    const float inv_tau = 1.0f / tau;
    // BF16X3: tile-parallel tcgen05 GEMMs (any N).  FP32: shared-memory path for N <= 64, FMA tiles beyond.
    // BF16X3: one-tile sizes run the shared-memory kernels with warp-level MMAs; beyond that the tile-parallel tcgen05 path
    if (precision == CRW_PREC_BF16X3) {
        if (walk_use_fused(B, T, N, C) && aligned16p(x)) return walk_fused_forward(x, B, T, N, C, tau, loss, A_or_null, saved, st);
        if (walk_small_mma_supported(N, C) && aligned16p(x) && !getenv("CRW_WALK_FORCE_TILES"))
            return walk_small_forward(x, B, T, N, C, tau, loss, A_or_null, ws, st, true);
        return walk_tiles_forward(x, B, T, N, C, tau, loss, A_or_null, ws, st);
    }
    if (walk_small_supported(N, C) && aligned16p(x)) return walk_small_forward(x, B, T, N, C, tau, loss, A_or_null, ws, st, false);
    walk_affinity_kernel<<<dim3(T - 1, B), kWT, 2 * N * sizeof(float), st>>>(x, ws, A_or_null, B, T, N, C, inv_tau);
    CRW_LAUNCH_RET();
    if (T < 3) {   // model.py:33-35: empty loop, loss = 0
        walk_zero_loss_kernel<<<1, 1, 0, st>>>(loss);
        CRW_LAUNCH_RET();
        return CRW_OK;
    }
    walk_chain_kernel<<<dim3(2, B), kWT, 0, st>>>(ws, B, T, N, C);
    CRW_LAUNCH_RET();
    walk_cycle_kernel<<<dim3(T - 2, B), kWT, 0, st>>>(ws, B, T, N, C);
    CRW_LAUNCH_RET();
    walk_loss_reduce_kernel<<<1, 32, 0, st>>>(ws, loss, B, T, N, C);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_walk_backward(const float* x, const void* saved, size_t saved_bytes, const float* dloss,
                                 const float* dA_or_null, int B, int T, int N, int C, float tau, int precision,
                                 float* dx, void* scratch, size_t scratch_bytes, void* stream) {
    if (!x || !saved || !dloss || !dx || B < 1 || T < 2 || N < 1 || C < 1 || !(tau > 0.0f)) return CRW_ERR_INVALID;
    if (precision != CRW_PREC_FP32 && precision != CRW_PREC_BF16X3) return CRW_ERR_UNSUPPORTED;
    if (saved_bytes < crw_walk_saved_bytes(B, T, N, C, precision)) return CRW_ERR_WORKSPACE;
    if (!scratch || scratch_bytes < crw_walk_backward_scratch_bytes(B, T, N, C, precision)) return CRW_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const float* ws = align256(const_cast<void*>(saved));
    float* sc = align256(scratch);
    const float inv_tau = 1.0f / tau;
    if (precision == CRW_PREC_BF16X3) {
        if (walk_use_fused(B, T, N, C) && aligned16p(x)) return walk_fused_backward(x, saved, dloss, dA_or_null, B, T, N, C, tau, dx, scratch, st);
        if (walk_small_mma_supported(N, C) && aligned16p(x) && !getenv("CRW_WALK_FORCE_TILES"))
            return walk_small_backward(x, ws, dloss, dA_or_null, B, T, N, C, tau, dx, sc, st, true);
        return walk_tiles_backward(x, ws, dloss, dA_or_null, B, T, N, C, tau, dx, sc, st);
    }
    if (walk_small_supported(N, C) && aligned16p(x))
        return walk_small_backward(x, ws, dloss, dA_or_null, B, T, N, C, tau, dx, sc, st, false);
    if (T >= 3) {
        walk_bwd_own_kernel<<<dim3(T - 2, B, 2), kWT, 0, st>>>(ws, sc, dloss, B, T, N, C);
        CRW_LAUNCH_RET();
        if (T >= 4) {
            walk_bwd_chain_kernel<<<dim3(2, B), kWT, 0, st>>>(ws, sc, B, T, N, C);
            CRW_LAUNCH_RET();
        }
    }
    walk_bwd_dA_kernel<<<dim3(T - 1, B), kWT, 2 * N * sizeof(float), st>>>(ws, sc, dA_or_null, B, T, N, C);
    CRW_LAUNCH_RET();
    walk_bwd_dx_kernel<<<dim3(T, B), kWT, 0, st>>>(x, ws, sc, dx, B, T, N, C, inv_tau);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
