// walk_tc_tiles.cu -- tile-parallel tensor-core walk (precision = CRW_PREC_BF16X3), any N.
//
// Same algorithm, workspace layout and reference mapping as walk_f32.cu (src/model.py:22-46 and its autograd), but
// every GEMM is a grid of independent 128 x 128 output tiles (one CTA each: tcgen05.mma bf16x3, TMEM accumulator,
// walk_tc.cuh) and the row-wise work (softmax, cycle cross-entropy, softmax backward, normalise backward) lives in
// separate small kernels.  The L / R chains and their adjoints are one launch per step (both chains in one grid),
// so the only serialisation left is the algorithm's own: 2(T-3) dependent products forward and backward.
#include "common.cuh"
#include "walk_layout.cuh"
#include "walk_tc.cuh"

namespace crw {

constexpr int kTT = 256;

template <class P>
__global__ void __launch_bounds__(kTT, 1) tc_tiles_kernel(P p) {
    extern __shared__ uint8_t tc_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t slot;
    TcGemmCtx cx;
    tc_ctx_init(cx, tc_raw, bars, &slot);
    p.run((int)blockIdx.z, (int)blockIdx.y * kTcTile, (int)blockIdx.x * kTcTile, cx);
    tc_ctx_fini(cx);
}

struct Dims { int B, T, N, C; };

// ---- forward problems -----------------------------------------------------------------------------------
struct AffinityProb {       // batch = b*(T-1) + t
    Dims d; const float* x; float* ws; float* A_out; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const int b = z / (d.T - 1), t = z % (d.T - 1), N = d.N;
        const float* x0 = x + ((size_t)b * d.T + t) * N * d.C;
        const float* i0 = ws + lay.invn + ((size_t)b * d.T + t) * N;
        float* At = ws + lay.mat(lay.A, b, t);
        float* Ao = A_out ? A_out + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
        const float it = inv_tau;
        cta_gemm_tc_tile<false, true>(x0, d.C, x0 + (size_t)N * d.C, d.C, N, N, d.C, nullptr, m0, n0, cx, [&](int m, int n, float v) {
            const float a = v * i0[m] * i0[N + n] * it;     // invn of frame t+1 follows frame t
            At[(size_t)m * N + n] = a;
            if (Ao) Ao[(size_t)m * N + n] = a;
        });
    }
};

struct ChainProb {          // batch = role*B + b ; step k
    Dims d; float* ws; int k;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const int role = z / d.B, b = z % d.B, N = d.N;
        if (role == 1 && k < 2) return;
        const float* A = role == 0 ? ws + lay.mat(lay.L, b, k - 1) : ws + lay.mat(lay.S, b, k - 1);
        const float* Bm = role == 0 ? ws + lay.mat(lay.Sp, b, k - 1) : ws + lay.mat(lay.R, b, k - 1);
        float* out = ws + lay.mat(role == 0 ? lay.L : lay.R, b, k);
        cta_gemm_tc_tile<false, false>(A, N, Bm, N, N, N, N, nullptr, m0, n0, cx, [&](int m, int n, float v) { out[(size_t)m * N + n] = v; });
    }
};

struct CycleProb {          // batch = b*K + (k-1)
    Dims d; float* ws;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const int K = d.T - 2, b = z / K, k = z % K + 1, N = d.N;
        float* G = ws + lay.mat(lay.G, b, k);
        cta_gemm_tc_tile<false, false>(ws + lay.mat(lay.L, b, k), N, ws + lay.mat(lay.R, b, k), N, N, N, N, nullptr, m0, n0, cx,
                                       [&](int m, int n, float v) { G[(size_t)m * N + n] = v; });
    }
};

// ---- backward problems ----------------------------------------------------------------------------------
struct OwnProb {            // batch = role*B*K + b*K + (k-1)
    Dims d; const float* ws; float* sc; const float* dloss;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const int K = d.T - 2, N = d.N, role = z / (d.B * K), r = z % (d.B * K), b = r / K, k = r % K + 1;
        const float s = *dloss / ((float)d.B * (float)N * (float)N);
        const float* G = ws + lay.mat(lay.G, b, k);
        if (role == 0) {
            float* o = sc + lay.mat(bl.dL, b, k);
            cta_gemm_tc_tile<false, true>(G, N, ws + lay.mat(lay.R, b, k), N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] = v * s; });
        } else {
            float* o = sc + lay.mat(bl.dR, b, k);
            cta_gemm_tc_tile<true, false>(ws + lay.mat(lay.L, b, k), N, G, N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] = v * s; });
        }
    }
};

struct BwdChainProb {       // batch = role*B + b ; step j: dL_j += dL_{j+1} S'_j^T ; dR_j += S_j^T dR_{j+1}
    Dims d; const float* ws; float* sc; int j;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const int role = z / d.B, b = z % d.B, N = d.N;
        if (role == 1 && j < 2) return;
        if (role == 0) {
            float* o = sc + lay.mat(bl.dL, b, j);
            cta_gemm_tc_tile<false, true>(sc + lay.mat(bl.dL, b, j + 1), N, ws + lay.mat(lay.Sp, b, j), N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] += v; });
        } else {
            float* o = sc + lay.mat(bl.dR, b, j);
            cta_gemm_tc_tile<true, false>(ws + lay.mat(lay.S, b, j), N, sc + lay.mat(bl.dR, b, j + 1), N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] += v; });
        }
    }
};

struct DsProb {             // batch = role*B*(T-1) + b*(T-1) + t ; role 0: dS'_t = L_t^T dL_{t+1} ; role 1: dS_t = dR_{t+1} R_t^T
    Dims d; const float* ws; float* sc;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const int K = d.T - 2, N = d.N, nt = d.T - 1, role = z / (d.B * nt), r = z % (d.B * nt), b = r / nt, t = r % nt;
        if (role == 0) {
            if (t + 1 > K) return;
            float* o = sc + lay.mat(bl.dSp, b, t);
            cta_gemm_tc_tile<true, false>(ws + lay.mat(lay.L, b, t), N, sc + lay.mat(bl.dL, b, t + 1), N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] = v; });
        } else {
            if (t < 1 || t + 1 > K) return;
            float* o = sc + lay.mat(bl.dS, b, t);
            cta_gemm_tc_tile<false, true>(sc + lay.mat(bl.dR, b, t + 1), N, ws + lay.mat(lay.R, b, t), N, N, N, N, nullptr, m0, n0, cx,
                                          [&](int m, int n, float v) { o[(size_t)m * N + n] = v; });
        }
    }
};

struct DxProb {             // batch = b*T + t ; dE_t = (dA_t E_{t+1} + dA_{t-1}^T E_{t-1}) / tau   (E = x * invn)
    Dims d; const float* x; const float* ws; const float* sc; float* dx; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const int b = z / d.T, t = z % d.T, N = d.N, C = d.C;
        float* o = dx + ((size_t)b * d.T + t) * N * C;
        const float* invn = ws + lay.invn + (size_t)b * d.T * N;
        const float it = inv_tau;
        if (t <= d.T - 2) {
            cta_gemm_tc_tile<false, false>(sc + lay.mat(bl.dAw, b, t), N, x + ((size_t)b * d.T + t + 1) * N * C, C, N, C, N,
                                           invn + (size_t)(t + 1) * N, m0, n0, cx,
                                           [&](int m, int c, float v) { o[(size_t)m * C + c] = v * it; });
        } else {
            for (int e = threadIdx.x; e < kTcTile * kTcTile; e += kTT) {
                const int m = m0 + e / kTcTile, c = n0 + e % kTcTile;
                if (m < N && c < C) o[(size_t)m * C + c] = 0.0f;
            }
            __syncthreads();
        }
        if (t >= 1)
            cta_gemm_tc_tile<true, false>(sc + lay.mat(bl.dAw, b, t - 1), N, x + ((size_t)b * d.T + t - 1) * N * C, C, N, C, N,
                                          invn + (size_t)(t - 1) * N, m0, n0, cx,
                                          [&](int m, int c, float v) { o[(size_t)m * C + c] += v * it; });
    }
};

// ---- row-wise kernels -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) t_rownorm_kernel(const float* __restrict__ x, float* ws, Dims d) {
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), rows = (long long)d.B * d.T * d.N;
    if (row >= rows) return;
    const float* xr = x + row * d.C;
    float ss = 0.0f;
    for (int c = lane; c < d.C; c += 32) ss = fmaf(xr[c], xr[c], ss);
    ss = warp_sum(ss);
    if (lane == 0) ws[lay.invn + row] = 1.0f / fmaxf(sqrtf(ss), kNormEps);
}

__global__ void __launch_bounds__(256) t_identity_kernel(float* ws, Dims d) {   // L_0 = I, R_1 = I
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int b = blockIdx.x, N = d.N;
    float* L0 = ws + lay.mat(lay.L, b, 0);
    float* R1 = (d.T >= 3) ? ws + lay.mat(lay.R, b, 1) : nullptr;
    for (size_t i = threadIdx.x; i < (size_t)N * N; i += blockDim.x) {
        const float v = (i / N == i % N) ? 1.0f : 0.0f;
        L0[i] = v;
        if (R1) R1[i] = v;
    }
}

__global__ void __launch_bounds__(256) t_softmax_kernel(float* ws, Dims d) {    // grid (T-1, B): S_t, S'_t from A_t
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* At = ws + lay.mat(lay.A, b, t);
    float* S = ws + lay.mat(lay.S, b, t);
    float* Sp = ws + lay.mat(lay.Sp, b, t);
    for (int r = warp; r < 2 * N; r += 8) {
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

__global__ void __launch_bounds__(256) t_cycle_epi_kernel(float* ws, Dims d) {  // grid (T-2, B): G_k = softmax(M_k) - I, loss partial
    __shared__ float red[8];
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int k = blockIdx.x + 1, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Gk = ws + lay.mat(lay.G, b, k);
    float part = 0.0f;
    for (int r = warp; r < N; r += 8) {
        float* row = Gk + (size_t)r * N;
        float mx = -INFINITY;
        for (int c = lane; c < N; c += 32) mx = fmaxf(mx, row[c]);
        mx = warp_max(mx);
        float se = 0.0f;
        for (int c = lane; c < N; c += 32) se += __expf(row[c] - mx);
        se = warp_sum(se);
        const float diag = row[r];
        __syncwarp();
        const float inv = 1.0f / se;
        for (int c = lane; c < N; c += 32) row[c] = __expf(row[c] - mx) * inv - (c == r ? 1.0f : 0.0f);
        part += (logf(se) + mx) - diag;
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w];
        ws[lay.part + (size_t)b * (d.T - 1) + k] = s;
    }
}

__global__ void t_loss_reduce_kernel(const float* ws, float* loss, Dims d) {
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int lane = threadIdx.x;
    float s = 0.0f;
    for (int i = lane; i < d.B * (d.T - 2); i += 32) s += ws[lay.part + (size_t)(i / (d.T - 2)) * (d.T - 1) + i % (d.T - 2) + 1];
    s = warp_sum(s);
    if (lane == 0) *loss = s / ((float)d.B * (float)d.N) / (float)d.N;
}
__global__ void t_zero_loss_kernel(float* loss) { *loss = 0.0f; }

__global__ void __launch_bounds__(256) t_dA_epi_kernel(const float* ws, float* sc, const float* dA_ext, Dims d) {   // grid (T-1, B)
    extern __shared__ float rdot[];   // rS[N], rSp[N]
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const BwdLayout bl(d.B, d.T, d.N);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, K = d.T - 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hasSp = (t + 1 <= K), hasS = (t >= 1 && t + 1 <= K);
    const float* S = ws + lay.mat(lay.S, b, t);
    const float* Sp = ws + lay.mat(lay.Sp, b, t);
    const float* dS = sc + lay.mat(bl.dS, b, t);
    const float* dSp = sc + lay.mat(bl.dSp, b, t);
    float* dA = sc + lay.mat(bl.dAw, b, t);
    const float* ext = dA_ext ? dA_ext + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
    for (int r = warp; r < 2 * N; r += 8) {
        const bool second = r >= N;
        const int i = second ? r - N : r;
        float a = 0.0f;
        if (second ? hasSp : hasS) {
            const float* P = (second ? Sp : S) + (size_t)i * N;
            const float* dP = (second ? dSp : dS) + (size_t)i * N;
            for (int j = lane; j < N; j += 32) a = fmaf(P[j], dP[j], a);
            a = warp_sum(a);
        }
        if (lane == 0) rdot[r] = a;
    }
    __syncthreads();
    for (size_t e = threadIdx.x; e < (size_t)N * N; e += blockDim.x) {
        const int i = (int)(e / N), j = (int)(e % N);
        float g = ext ? ext[e] : 0.0f;
        if (hasS) g += S[e] * (dS[e] - rdot[i]);
        if (hasSp) g += Sp[(size_t)j * N + i] * (dSp[(size_t)j * N + i] - rdot[N + j]);
        dA[e] = g;
    }
}

__global__ void __launch_bounds__(256) t_dx_epi_kernel(const float* __restrict__ x, const float* ws, float* dx, Dims d) {   // grid (T, B)
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, C = d.C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* invn = ws + lay.invn + ((size_t)b * d.T + t) * N;
    const float* xt = x + ((size_t)b * d.T + t) * N * C;
    float* o = dx + ((size_t)b * d.T + t) * N * C;
    for (int i = warp; i < N; i += 8) {
        const float inv = invn[i];
        float* orow = o + (size_t)i * C;
        const float* xr = xt + (size_t)i * C;
        if (inv >= 1.0f / kNormEps) {
            for (int c = lane; c < C; c += 32) orow[c] *= inv;
            continue;
        }
        float dot = 0.0f;
        for (int c = lane; c < C; c += 32) dot = fmaf(xr[c] * inv, orow[c], dot);
        dot = warp_sum(dot);
        for (int c = lane; c < C; c += 32) orow[c] = (orow[c] - xr[c] * inv * dot) * inv;
    }
}

// ---- host orchestration -----------------------------------------------------------------------------------
template <class P>
static int launch_tiles(const P& p, int Mrows, int Ncols, int batch, cudaStream_t st) {
    static bool opted = false;   // per instantiation; idempotent
    if (!opted) {
        CRW_CUDA_RET(cudaFuncSetAttribute(tc_tiles_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        opted = true;
    }
    if (batch <= 0) return CRW_OK;
    dim3 grid(ceil_div(Ncols, kTcTile), ceil_div(Mrows, kTcTile), batch);
    tc_tiles_kernel<P><<<grid, kTT, kTcSmemBytes, st>>>(p);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// No extra state beyond the fp32 workspaces: operands are converted to bf16 hi/lo while they are staged.
size_t walk_tiles_saved_extra_bytes(int, int, int, int) { return 0; }
size_t walk_tiles_scratch_extra_bytes(int, int, int, int) { return 0; }

int walk_tiles_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st) {
    const Dims d{B, T, N, C};
    const long long rows = (long long)B * T * N;
    t_rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, ws, d);
    CRW_LAUNCH_RET();
    int rc = launch_tiles(AffinityProb{d, x, ws, A_or_null, 1.0f / tau}, N, N, B * (T - 1), st);
    if (rc) return rc;
    if (T < 3) {
        t_zero_loss_kernel<<<1, 1, 0, st>>>(loss);
        CRW_LAUNCH_RET();
        return CRW_OK;
    }
    t_softmax_kernel<<<dim3(T - 1, B), 256, 0, st>>>(ws, d);
    CRW_LAUNCH_RET();
    t_identity_kernel<<<B, 256, 0, st>>>(ws, d);
    CRW_LAUNCH_RET();
    const int K = T - 2;
    for (int k = 1; k <= K; ++k)
        if ((rc = launch_tiles(ChainProb{d, ws, k}, N, N, k >= 2 ? 2 * B : B, st))) return rc;
    if ((rc = launch_tiles(CycleProb{d, ws}, N, N, B * K, st))) return rc;
    t_cycle_epi_kernel<<<dim3(K, B), 256, 0, st>>>(ws, d);
    CRW_LAUNCH_RET();
    t_loss_reduce_kernel<<<1, 32, 0, st>>>(ws, loss, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

int walk_tiles_backward(const float* x, const float* ws, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st) {
    const Dims d{B, T, N, C};
    const int K = T - 2;
    int rc;
    if (T >= 3) {
        if ((rc = launch_tiles(OwnProb{d, ws, sc, dloss}, N, N, 2 * B * K, st))) return rc;
        for (int j = K - 1; j >= 1; --j)
            if ((rc = launch_tiles(BwdChainProb{d, ws, sc, j}, N, N, j >= 2 ? 2 * B : B, st))) return rc;
        if ((rc = launch_tiles(DsProb{d, ws, sc}, N, N, 2 * B * (T - 1), st))) return rc;
    }
    t_dA_epi_kernel<<<dim3(T - 1, B), 256, 2 * N * sizeof(float), st>>>(ws, sc, dA_or_null, d);
    CRW_LAUNCH_RET();
    if ((rc = launch_tiles(DxProb{d, x, ws, sc, dx, 1.0f / tau}, N, C, B * T, st))) return rc;
    t_dx_epi_kernel<<<dim3(T, B), 256, 0, st>>>(x, ws, dx, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

}  // namespace crw
