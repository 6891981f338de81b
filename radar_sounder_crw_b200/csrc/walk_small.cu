// walk_small.cu -- shared-memory-resident fp32 walk kernels for the reference's own sizes (N <= 64, C <= 128).
//
// Same algorithm, workspace layout and reference mapping as walk_f32.cu (src/model.py:22-46 and its autograd);
// the difference is where the operands live.  At N = 47..49 every stage is latency bound, so each CTA pulls its
// operands into shared memory ONCE with cp.async (zero-padded to 64 x 64), runs the GEMM from shared memory with
// no barrier inside the k-loop, keeps intermediates (A_t, M_k, the running chain product) on chip, and the two
// sequential chains prefetch the next step's operand while the current product is being formed.
#include "common.cuh"
#include "walk_layout.cuh"

namespace crw {

constexpr int kST = 256;      // threads per CTA (16 x 16 grid of 4x4 register blocks = one 64 x 64 tile)
constexpr int kLD = 68;       // row pitch of a 64 x 64 matrix in smem (floats)
constexpr int kLDX = 132;     // row pitch of a 64 x 128 feature tile in smem
constexpr int kMat = 64 * kLD;
constexpr int kMatX = 64 * kLDX;

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void zero_fill(float* p, int n) {     // n % 4 == 0, p 16-byte aligned
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int i = threadIdx.x; i < (n >> 2); i += kST) p4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}
// N x N matrix (global, dense) -> smem pitch kLD.  Rows are not 16-byte aligned for odd N: 4-byte cp.async.
// (x / d for x * d < 2^32 is __umulhi(x, 2^32 / d + 1): the index loops below would otherwise spend their time dividing)
__device__ __forceinline__ unsigned div_magic(int d) { return (unsigned)(0x100000000ull / (unsigned)d) + 1u; }
__device__ __forceinline__ void load_nn(float* dst, const float* src, int N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < N; r += kST / 32) {           // N <= 64: at most two columns per lane
        if (lane < N) cp_async4(dst + r * kLD + lane, src + r * N + lane);
        if (lane + 32 < N) cp_async4(dst + r * kLD + lane + 32, src + r * N + lane + 32);
    }
}
// N x C feature tile (global rows are 16-byte aligned since C % 4 == 0) -> smem pitch kLDX
__device__ __forceinline__ void load_nc(float* dst, const float* src, int N, int C) {
    const int c4 = C >> 2;
    const unsigned mg = div_magic(c4);
    for (int i = threadIdx.x; i < N * c4; i += kST) {
        const int r = (int)__umulhi((unsigned)i, mg), c = (i - r * c4) * 4;
        cp_async16(dst + r * kLDX + c, src + (size_t)r * C + c);
    }
}

// acc[i][j] += sum_k opA[row(i)][k] * opB[k][col(j)], all operands in smem, K4 = K rounded up to 4 (padding is zero).
//   !TA: A[m*lda + k], rows owned: ty + 16 i      TA: A[k*lda + m], rows owned: 4 ty + i
//   !TB: B[k*ldb + n], cols owned: 4 tx + j        TB: B[n*ldb + k], cols owned: tx + 16 j
template <bool TA> __device__ __forceinline__ int own_row(int i) { return TA ? (threadIdx.x >> 4) * 4 + i : (threadIdx.x >> 4) + 16 * i; }
template <bool TB> __device__ __forceinline__ int own_col(int j) { return TB ? (threadIdx.x & 15) + 16 * j : (threadIdx.x & 15) * 4 + j; }

template <bool TA, bool TB>
__device__ __forceinline__ void sgemm_tile(const float* A, int lda, const float* B, int ldb, int K4, int n0, float (&acc)[4][4]) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll 2
    for (int k = 0; k < K4; k += 4) {
        float a[4][4], b[4][4];   // a[i][kk], b[kk][j]
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!TA) {
                const float4 v = *reinterpret_cast<const float4*>(A + (ty + 16 * u) * lda + k);
                a[u][0] = v.x; a[u][1] = v.y; a[u][2] = v.z; a[u][3] = v.w;
            } else {
                const float4 v = *reinterpret_cast<const float4*>(A + (k + u) * lda + ty * 4);
                a[0][u] = v.x; a[1][u] = v.y; a[2][u] = v.z; a[3][u] = v.w;
            }
            if (!TB) {
                const float4 v = *reinterpret_cast<const float4*>(B + (k + u) * ldb + n0 + tx * 4);
                b[u][0] = v.x; b[u][1] = v.y; b[u][2] = v.z; b[u][3] = v.w;
            } else {
                const float4 v = *reinterpret_cast<const float4*>(B + (n0 + tx + 16 * u) * ldb + k);
                b[0][u] = v.x; b[1][u] = v.y; b[2][u] = v.z; b[3][u] = v.w;
            }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[kk][j], acc[i][j]);
    }
}
__device__ __forceinline__ void zero_acc(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
}

// ---- tensor-core variant of the same 64 x 64 tile (precision = CRW_PREC_BF16X3 at the reference's sizes) ----
// mma.sync m16n8k16 on error-compensated bf16 pairs split on the fly from the fp32 operands in shared memory
// (x = hi + lo; hi.hi + hi.lo + lo.hi, fp32 accumulate).  A dependent chain of 47 x 47 products is latency bound: the
// warp-level MMA reads the operands where the previous step left them (no descriptor / TMEM round trip per step), which
// is why this path does not use tcgen05 -- walk_tc_tiles.cu does, for N beyond one tile.
// Warp w owns rows 16 (w & 3) .. +15 and columns 32 (w >> 2) .. +31; acc[nt][r] = n8-tile nt, fragment register r.
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));      // low half <- x0
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// same operand conventions as sgemm_tile; K4 is rounded up to 16 here (the padding of every operand is zero or unused)
template <bool TA, bool TB>
__device__ __forceinline__ void mma_tile(const float* A, int lda, const float* B, int ldb, int K4, int n0, float (&acc)[4][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int m0 = 16 * (warp & 3), c0 = n0 + 32 * (warp >> 2);
    for (int k0 = 0; k0 < K4; k0 += 16) {
        uint32_t ah[4], al[4];                 // a0: (g, k..), a1: (g+8, k..), a2: (g, k+8..), a3: (g+8, k+8..)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = m0 + g + 8 * r, k = k0 + 2 * tig + 8 * h;
                float x0, x1;
                if (!TA) { const float2 v = *reinterpret_cast<const float2*>(A + m * lda + k); x0 = v.x; x1 = v.y; }
                else { x0 = A[k * lda + m]; x1 = A[(k + 1) * lda + m]; }
                split_pair(x0, x1, ah[2 * h + r], al[2 * h + r]);
            }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int n = c0 + 8 * nt + g;
            uint32_t bh[2], bl[2];             // b0: (k.., n), b1: (k+8.., n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + 2 * tig + 8 * h;
                float x0, x1;
                if (TB) { const float2 v = *reinterpret_cast<const float2*>(B + n * ldb + k); x0 = v.x; x1 = v.y; }
                else { x0 = B[k * ldb + n]; x1 = B[(k + 1) * ldb + n]; }
                split_pair(x0, x1, bh[h], bl[h]);
            }
            mma_bf16(acc[nt], ah, bh[0], bh[1]);
            mma_bf16(acc[nt], ah, bl[0], bl[1]);
            mma_bf16(acc[nt], al, bh[0], bh[1]);
        }
    }
}
template <bool MMA, bool TA, bool TB>
__device__ __forceinline__ void gemm_tile(const float* A, int lda, const float* B, int ldb, int K4, int n0, float (&acc)[4][4]) {
    if (MMA) mma_tile<TA, TB>(A, lda, B, ldb, K4, n0, acc);
    else sgemm_tile<TA, TB>(A, lda, B, ldb, K4, n0, acc);
}
// (row, column) of acc[i][j] within the 64 x 64 tile
template <bool MMA, bool TA, bool TB>
__device__ __forceinline__ void out_rc(int i, int j, int& m, int& n) {
    if (MMA) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        m = 16 * (warp & 3) + (lane >> 2) + 8 * (j >> 1);
        n = 32 * (warp >> 2) + 8 * i + 2 * (lane & 3) + (j & 1);
    } else {
        m = own_row<TA>(i);
        n = own_col<TB>(j);
    }
}

// ------------------------------------------------------------------------------------------
// forward 1: grid (T-1, B).  A_t = E_t E_{t+1}^T / tau, S_t, S'_t   (model.py:22,26 + the softmaxes of :44)
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_affinity_kernel(const float* __restrict__ x, float* ws, float* A_out, int B, int T,
                                                             int N, int C, float inv_tau) {
    extern __shared__ __align__(16) float sm[];
    float* X0 = sm;
    float* X1 = X0 + kMatX;
    float* At = X1 + kMatX;
    float* inv0 = At + kMat;
    float* inv1 = inv0 + 64;
    const WalkLayout lay(B, T, N, C);
    const int t = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* x0 = x + ((size_t)b * T + t) * N * C;
    load_nc(X0, x0, N, C);
    load_nc(X1, x0 + (size_t)N * C, N, C);
    cp_async_wait_all();
    __syncthreads();
    for (int r = warp; r < 2 * N; r += kST / 32) {
        const float* xr = (r < N) ? X0 + r * kLDX : X1 + (r - N) * kLDX;
        float ss = 0.0f;
        for (int c = 4 * lane; c < C; c += 128) {          // C % 4 == 0, rows 16-byte aligned
            const float4 v = *reinterpret_cast<const float4*>(xr + c);
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
        if (lane == 0) {
            inv0[r < N ? r : 64 + (r - N)] = inv;
            if (r < N) ws[lay.invn + ((size_t)b * T + t) * N + r] = inv;
            else if (t == T - 2) ws[lay.invn + ((size_t)b * T + t + 1) * N + (r - N)] = inv;
        }
    }
    __syncthreads();
    float acc[4][4];
    zero_acc(acc);
    gemm_tile<MMA, false, true>(X0, kLDX, X1, kLDX, C, 0, acc);
    float* Ao = A_out ? A_out + ((size_t)b * (T - 1) + t) * N * N : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m, n;
            out_rc<MMA, false, true>(i, j, m, n);
            if (m < N && n < N) {
                const float a = acc[i][j] * inv0[m] * inv1[n] * inv_tau;
                At[m * kLD + n] = a;
                if (Ao) Ao[(size_t)m * N + n] = a;
            }
        }
    __syncthreads();
    float* S = ws + lay.mat(lay.S, b, t);
    float* Sp = ws + lay.mat(lay.Sp, b, t);
    for (int r = warp; r < 2 * N; r += kST / 32) {
        const bool col = r >= N;
        const int i = col ? r - N : r;
        const int step = col ? kLD : 1, base = col ? i : i * kLD;
        // N <= 64: a lane holds its (at most) two entries of the row / column in registers across the three passes
        const bool has1 = lane + 32 < N;
        const float a0 = (lane < N) ? At[base + lane * step] : -INFINITY;
        const float a1 = has1 ? At[base + (lane + 32) * step] : -INFINITY;
        const float mx = warp_max(fmaxf(a0, a1));
        const float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx);          // exp(-inf) = 0 for the absent entries
        const float inv = 1.0f / warp_sum(e0 + e1);
        float* dst = (col ? Sp : S) + (size_t)i * N;
        if (lane < N) dst[lane] = e0 * inv;
        if (has1) dst[lane + 32] = e1 * inv;
    }
}

// store the thread's 4x4 block of a 64 x 64 result into smem (pitch kLD) and/or a dense N x N global matrix
template <bool MMA, bool TA, bool TB, class F>
__device__ __forceinline__ void for_each_out(const float (&acc)[4][4], int N, F f) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m, n;
            out_rc<MMA, TA, TB>(i, j, m, n);
            if (m < N && n < N) f(m, n, acc[i][j]);
        }
}

// ------------------------------------------------------------------------------------------
// forward 2: grid (2, B).  x = 0: L_k = L_{k-1} S'_{k-1};  x = 1: R_k = S_{k-1} R_{k-1}   (sequential in k)
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_chain_kernel(float* ws, int B, int T, int N, int C) {
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                 // running product
    float* Op = P + kMat;          // [2] next operand
    const WalkLayout lay(B, T, N, C);
    const int b = blockIdx.y, K = T - 2, N4 = (N + 3) & ~3;
    const bool isL = blockIdx.x == 0;
    const size_t chain = isL ? lay.L : lay.R, opnd = isL ? lay.Sp : lay.S;
    const int k_first = isL ? 1 : 2;
    zero_fill(sm, 3 * kMat);
    __syncthreads();
    if (k_first <= K) load_nn(Op, ws + lay.mat(opnd, b, k_first - 1), N);
    float* first = ws + lay.mat(chain, b, k_first - 1);    // L_0 = I, R_1 = I
    const unsigned mgN = div_magic(N);
    for (int i = threadIdx.x; i < N * N; i += kST) {
        const int r = (int)__umulhi((unsigned)i, mgN), c = i - r * N;
        const float v = (r == c) ? 1.0f : 0.0f;
        P[r * kLD + c] = v;
        first[i] = v;
    }
    for (int k = k_first; k <= K; ++k) {
        const float* cur = Op + ((k - k_first) & 1) * kMat;
        cp_async_wait_all();
        __syncthreads();
        if (k + 1 <= K) load_nn(Op + ((k + 1 - k_first) & 1) * kMat, ws + lay.mat(opnd, b, k), N);
        float acc[4][4];
        zero_acc(acc);
        if (isL) gemm_tile<MMA, false, false>(P, kLD, cur, kLD, N4, 0, acc);
        else gemm_tile<MMA, false, false>(cur, kLD, P, kLD, N4, 0, acc);
        __syncthreads();
        float* out = ws + lay.mat(chain, b, k);
        for_each_out<MMA, false, false>(acc, N, [&](int m, int n, float v) { P[m * kLD + n] = v; out[(size_t)m * N + n] = v; });
    }
}

// ------------------------------------------------------------------------------------------
// forward 3: grid (T-2, B).  M_k = L_k R_k; loss partial; G_k = rowsoftmax(M_k) - I   (model.py:45)
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_cycle_kernel(float* ws, int B, int T, int N, int C) {
    extern __shared__ __align__(16) float sm[];
    float* Lk = sm;
    float* Rk = Lk + kMat;
    float* M = Rk + kMat;
    __shared__ float red[kST / 32];
    const WalkLayout lay(B, T, N, C);
    const int k = blockIdx.x + 1, b = blockIdx.y, N4 = (N + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    zero_fill(sm, 2 * kMat);
    __syncthreads();
    load_nn(Lk, ws + lay.mat(lay.L, b, k), N);
    load_nn(Rk, ws + lay.mat(lay.R, b, k), N);
    cp_async_wait_all();
    __syncthreads();
    float acc[4][4];
    zero_acc(acc);
    gemm_tile<MMA, false, false>(Lk, kLD, Rk, kLD, N4, 0, acc);
    for_each_out<MMA, false, false>(acc, N, [&](int m, int n, float v) { M[m * kLD + n] = v; });
    __syncthreads();
    float* Gk = ws + lay.mat(lay.G, b, k);
    float part = 0.0f;
    for (int d = warp; d < N; d += kST / 32) {
        const float* row = M + d * kLD;
        const bool has1 = lane + 32 < N;
        const float a0 = (lane < N) ? row[lane] : -INFINITY, a1 = has1 ? row[lane + 32] : -INFINITY;
        const float mx = warp_max(fmaxf(a0, a1));
        const float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx);
        const float se = warp_sum(e0 + e1);
        const float inv = 1.0f / se;
        if (lane < N) Gk[(size_t)d * N + lane] = e0 * inv - (lane == d ? 1.0f : 0.0f);
        if (has1) Gk[(size_t)d * N + lane + 32] = e1 * inv - (lane + 32 == d ? 1.0f : 0.0f);
        part += (logf(se) + mx) - row[d];
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < kST / 32; ++w) s += red[w];
        ws[lay.part + (size_t)b * (T - 1) + k] = s;
    }
}

// ------------------------------------------------------------------------------------------
// backward 1: grid (T-2, B, 2).  z = 0: dL_k = s G_k R_k^T;  z = 1: dR_k = s L_k^T G_k
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_bwd_own_kernel(const float* ws, float* sc, const float* dloss, int B, int T, int N,
                                                            int C) {
    extern __shared__ __align__(16) float sm[];
    float* G = sm;
    float* O = G + kMat;
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int k = blockIdx.x + 1, b = blockIdx.y, N4 = (N + 3) & ~3;
    const float s = *dloss / ((float)B * (float)N * (float)N);
    zero_fill(sm, 2 * kMat);
    __syncthreads();
    load_nn(G, ws + lay.mat(lay.G, b, k), N);
    load_nn(O, ws + lay.mat(blockIdx.z == 0 ? lay.R : lay.L, b, k), N);
    cp_async_wait_all();
    __syncthreads();
    float acc[4][4];
    zero_acc(acc);
    if (blockIdx.z == 0) {
        gemm_tile<MMA, false, true>(G, kLD, O, kLD, N4, 0, acc);
        float* o = sc + lay.mat(bl.dL, b, k);
        for_each_out<MMA, false, true>(acc, N, [&](int m, int n, float v) { o[(size_t)m * N + n] = v * s; });
    } else {
        gemm_tile<MMA, true, false>(O, kLD, G, kLD, N4, 0, acc);
        float* o = sc + lay.mat(bl.dR, b, k);
        for_each_out<MMA, true, false>(acc, N, [&](int m, int n, float v) { o[(size_t)m * N + n] = v * s; });
    }
}

// ------------------------------------------------------------------------------------------
// backward 2: grid (2, B).  x = 0: dL_j += dL_{j+1} S'_j^T (j = K-1..1);  x = 1: dR_j += S_j^T dR_{j+1} (j = K-1..2)
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_bwd_chain_kernel(const float* ws, float* sc, int B, int T, int N, int C) {
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                  // running adjoint dL_{j+1} / dR_{j+1}
    float* Op = P + kMat;           // [2] S'_j or S_j
    float* Own = Op + 2 * kMat;     // [2] own term of step j
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int b = blockIdx.y, K = T - 2, N4 = (N + 3) & ~3;
    const bool isL = blockIdx.x == 0;
    const size_t adj = isL ? bl.dL : bl.dR, opnd = isL ? lay.Sp : lay.S;
    const int j_last = isL ? 1 : 2;
    zero_fill(sm, 5 * kMat);
    __syncthreads();
    load_nn(P, sc + lay.mat(adj, b, K), N);
    if (K - 1 >= j_last) {
        load_nn(Op, ws + lay.mat(opnd, b, K - 1), N);
        load_nn(Own, sc + lay.mat(adj, b, K - 1), N);
    }
    for (int j = K - 1, it = 0; j >= j_last; --j, ++it) {
        const float* cur = Op + (it & 1) * kMat;
        const float* own = Own + (it & 1) * kMat;
        cp_async_wait_all();
        __syncthreads();
        if (j - 1 >= j_last) {
            load_nn(Op + ((it + 1) & 1) * kMat, ws + lay.mat(opnd, b, j - 1), N);
            load_nn(Own + ((it + 1) & 1) * kMat, sc + lay.mat(adj, b, j - 1), N);
        }
        float acc[4][4];
        zero_acc(acc);
        float* out = sc + lay.mat(adj, b, j);
        if (isL) {
            gemm_tile<MMA, false, true>(P, kLD, cur, kLD, N4, 0, acc);
            __syncthreads();
            for_each_out<MMA, false, true>(acc, N, [&](int m, int n, float v) {
                const float r = v + own[m * kLD + n];
                P[m * kLD + n] = r;
                out[(size_t)m * N + n] = r;
            });
        } else {
            gemm_tile<MMA, true, false>(cur, kLD, P, kLD, N4, 0, acc);
            __syncthreads();
            for_each_out<MMA, true, false>(acc, N, [&](int m, int n, float v) {
                const float r = v + own[m * kLD + n];
                P[m * kLD + n] = r;
                out[(size_t)m * N + n] = r;
            });
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward 3: grid (T-1, B).  dS'_t = L_t^T dL_{t+1};  dS_t = dR_{t+1} R_t^T;  softmax backward;  dA_t
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_bwd_dA_kernel(const float* ws, float* sc, const float* dA_ext, int B, int T, int N,
                                                           int C) {
    extern __shared__ __align__(16) float sm[];
    float* Lt = sm;
    float* dLn = Lt + kMat;
    float* Rt = dLn + kMat;
    float* dRn = Rt + kMat;
    float* Ss = dRn + kMat;
    float* Sps = Ss + kMat;
    float* dSs = Rt;               // results overwrite an operand that is dead by then (6 buffers -> 2 CTAs per SM)
    float* dSps = Lt;
    float* rS = Sps + kMat;
    float* rSp = rS + 64;
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const int t = blockIdx.x, b = blockIdx.y, K = T - 2, N4 = (N + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hasSp = (t + 1 <= K), hasS = (t >= 1 && t + 1 <= K);
    float* dA = sc + lay.mat(bl.dAw, b, t);
    const float* ext = dA_ext ? dA_ext + ((size_t)b * (T - 1) + t) * N * N : nullptr;
    if (!hasSp && !hasS) {
        for (int e = threadIdx.x; e < N * N; e += kST) dA[e] = ext ? ext[e] : 0.0f;
        return;
    }
    zero_fill(sm, 6 * kMat);
    __syncthreads();
    if (hasSp) {
        load_nn(Lt, ws + lay.mat(lay.L, b, t), N);
        load_nn(dLn, sc + lay.mat(bl.dL, b, t + 1), N);
        load_nn(Sps, ws + lay.mat(lay.Sp, b, t), N);
    }
    if (hasS) {
        load_nn(Rt, ws + lay.mat(lay.R, b, t), N);
        load_nn(dRn, sc + lay.mat(bl.dR, b, t + 1), N);
        load_nn(Ss, ws + lay.mat(lay.S, b, t), N);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[4][4], acc2[4][4];
    zero_acc(acc);
    zero_acc(acc2);
    if (hasSp) gemm_tile<MMA, true, false>(Lt, kLD, dLn, kLD, N4, 0, acc);
    if (hasS) gemm_tile<MMA, false, true>(dRn, kLD, Rt, kLD, N4, 0, acc2);
    __syncthreads();               // every read of Lt / Rt is done: overwrite them with the products
    for_each_out<MMA, true, false>(acc, N, [&](int m, int n, float v) { dSps[m * kLD + n] = v; });
    for_each_out<MMA, false, true>(acc2, N, [&](int m, int n, float v) { dSs[m * kLD + n] = v; });
    __syncthreads();
    for (int r = warp; r < 2 * N; r += kST / 32) {
        const bool second = r >= N;
        const int i = second ? r - N : r;
        const float* P = (second ? Sps : Ss) + i * kLD;
        const float* dP = (second ? dSps : dSs) + i * kLD;
        float a = (lane < N) ? P[lane] * dP[lane] : 0.0f;                  // zero when the term is absent (buffers zero-filled)
        if (lane + 32 < N) a = fmaf(P[lane + 32], dP[lane + 32], a);
        a = warp_sum(a);
        if (lane == 0) (second ? rSp : rS)[i] = a;
    }
    __syncthreads();
    const unsigned mgN = div_magic(N);
    for (int e = threadIdx.x; e < N * N; e += kST) {
        const int i = (int)__umulhi((unsigned)e, mgN), j = e - i * N;
        float g = ext ? ext[e] : 0.0f;
        g += Ss[i * kLD + j] * (dSs[i * kLD + j] - rS[i]);
        g += Sps[j * kLD + i] * (dSps[j * kLD + i] - rSp[j]);
        dA[e] = g;
    }
}

// ------------------------------------------------------------------------------------------
// backward 4: grid (T, B).  dE_t = (dA_t E_{t+1} + dA_{t-1}^T E_{t-1}) / tau, then the normalise backward
// ------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(kST) walk_s_bwd_dx_kernel(const float* __restrict__ x, const float* ws, const float* sc, float* dx,
                                                           int B, int T, int N, int C, float inv_tau) {
    extern __shared__ __align__(16) float sm[];
    float* dAt = sm;
    float* dAp = dAt + kMat;
    float* En = dAp + kMat;
    float* Ep = En + kMatX;
    float* Out = En;               // written only after both products are in registers (4 buffers -> 2 CTAs per SM)
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    // 1-D grid of B * T items, the B * (T - 2) inner frames first: they form two products each, frames 0 and T - 1 only one.
    // B * T = 320 CTAs at config 2 are 1.08 waves of two CTAs per SM; with the light items last the second wave is short
    // (and starts as soon as the first light items finish) instead of costing a whole heavy CTA time.
    int t, b;
    {
        const int inner = B * (T - 2), idx = blockIdx.x;
        if (idx < inner) { b = idx / (T - 2); t = 1 + idx - b * (T - 2); }
        else { const int j = idx - inner; b = j >> 1; t = (j & 1) ? T - 1 : 0; }
    }
    const int N4 = (N + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* invn = ws + lay.invn + (size_t)b * T * N;
    const bool hasN = t <= T - 2, hasP = t >= 1;
    zero_fill(sm, 2 * kMat + 2 * kMatX);
    __syncthreads();
    if (hasN) {
        load_nn(dAt, sc + lay.mat(bl.dAw, b, t), N);
        load_nc(En, x + ((size_t)b * T + t + 1) * N * C, N, C);
    }
    if (hasP) {
        load_nn(dAp, sc + lay.mat(bl.dAw, b, t - 1), N);
        load_nc(Ep, x + ((size_t)b * T + t - 1) * N * C, N, C);
    }
    cp_async_wait_all();
    __syncthreads();
    // E = x * invn (rows)
    for (int r = warp; r < N; r += kST / 32) {
        const float in = hasN ? invn[(size_t)(t + 1) * N + r] : 0.0f, ip = hasP ? invn[(size_t)(t - 1) * N + r] : 0.0f;
        for (int c = lane; c < C; c += 32) {
            if (hasN) En[r * kLDX + c] *= in;
            if (hasP) Ep[r * kLDX + c] *= ip;
        }
    }
    __syncthreads();
    float acc[2][4][4], acc2[2][4][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        zero_acc(acc[h]);
        zero_acc(acc2[h]);
        if (h * 64 < C) {
            if (hasN) gemm_tile<MMA, false, false>(dAt, kLD, En, kLDX, N4, h * 64, acc[h]);
            if (hasP) gemm_tile<MMA, true, false>(dAp, kLD, Ep, kLDX, N4, h * 64, acc2[h]);
        }
    }
    __syncthreads();               // En is dead: reuse it as the output tile
    // the two products own different rows per thread (row mapping depends on TA): combine through smem
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int m, n;
                out_rc<MMA, false, false>(i, j, m, n);
                n += h * 64;
                if (m < N && n < C) Out[m * kLDX + n] = acc[h][i][j] * inv_tau;
            }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int m, n;
                out_rc<MMA, true, false>(i, j, m, n);
                n += h * 64;
                if (m < N && n < C) Out[m * kLDX + n] += acc2[h][i][j] * inv_tau;
            }
    __syncthreads();
    const float* xt = x + ((size_t)b * T + t) * N * C;
    float* o = dx + ((size_t)b * T + t) * N * C;
    for (int i = warp; i < N; i += kST / 32) {
        const float inv = invn[(size_t)t * N + i];
        const float* dr = Out + i * kLDX;
        const float* xr = xt + (size_t)i * C;
        float* orow = o + (size_t)i * C;
        if (inv >= 1.0f / kNormEps) {   // ||x|| <= eps: F.normalize divides by the constant eps
            for (int c = lane; c < C; c += 32) orow[c] = dr[c] * inv;
            continue;
        }
        float dot = 0.0f;
        for (int c = lane; c < C; c += 32) dot = fmaf(xr[c] * inv, dr[c], dot);
        dot = warp_sum(dot);
        for (int c = lane; c < C; c += 32) orow[c] = (dr[c] - xr[c] * inv * dot) * inv;
    }
}

__global__ void walk_s_loss_reduce_kernel(const float* ws, float* loss, int B, int T, int N, int C) {
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
__global__ void walk_s_zero_loss_kernel(float* loss) { *loss = 0.0f; }

// opt in to > 48 KB dynamic shared memory (idempotent; cheap enough to repeat per launch, never inside the kernels)
template <class Kern>
static int set_smem(Kern kern, size_t bytes) {
    CRW_CUDA_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // one carveout for every kernel of the sequence: a different L1/shared split per launch forces an SM drain between them
    CRW_CUDA_RET(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return CRW_OK;
}

bool walk_small_supported(int N, int C) { return N <= 64 && C <= 128 && (C % 4) == 0; }
bool walk_small_mma_supported(int N, int C) { return walk_small_supported(N, C) && (C % 16) == 0; }

template <bool MMA>
static int walk_small_forward_t(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                                cudaStream_t st) {
    const float inv_tau = 1.0f / tau;
    const size_t sm1 = (2 * kMatX + kMat + 128) * sizeof(float), sm2 = 3 * kMat * sizeof(float);
    int rc = set_smem(walk_s_affinity_kernel<MMA>, sm1);
    if (rc) return rc;
    walk_s_affinity_kernel<MMA><<<dim3(T - 1, B), kST, sm1, st>>>(x, ws, A_or_null, B, T, N, C, inv_tau);
    CRW_LAUNCH_RET();
    if (T < 3) {
        walk_s_zero_loss_kernel<<<1, 1, 0, st>>>(loss);
        CRW_LAUNCH_RET();
        return CRW_OK;
    }
    if ((rc = set_smem(walk_s_chain_kernel<MMA>, sm2))) return rc;
    walk_s_chain_kernel<MMA><<<dim3(2, B), kST, sm2, st>>>(ws, B, T, N, C);
    CRW_LAUNCH_RET();
    if ((rc = set_smem(walk_s_cycle_kernel<MMA>, sm2))) return rc;
    walk_s_cycle_kernel<MMA><<<dim3(T - 2, B), kST, sm2, st>>>(ws, B, T, N, C);
    CRW_LAUNCH_RET();
    if ((rc = set_smem(walk_s_loss_reduce_kernel, 0))) return rc;
    walk_s_loss_reduce_kernel<<<1, 32, 0, st>>>(ws, loss, B, T, N, C);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

template <bool MMA>
static int walk_small_backward_t(const float* x, const float* ws, const float* dloss, const float* dA_or_null, int B, int T, int N,
                                 int C, float tau, float* dx, float* sc, cudaStream_t st) {
    const float inv_tau = 1.0f / tau;
    int rc;
    if (T >= 3) {
        const size_t sm = 2 * kMat * sizeof(float);
        if ((rc = set_smem(walk_s_bwd_own_kernel<MMA>, sm))) return rc;
        walk_s_bwd_own_kernel<MMA><<<dim3(T - 2, B, 2), kST, sm, st>>>(ws, sc, dloss, B, T, N, C);
        CRW_LAUNCH_RET();
        if (T >= 4) {
            const size_t smc = 5 * kMat * sizeof(float);
            if ((rc = set_smem(walk_s_bwd_chain_kernel<MMA>, smc))) return rc;
            walk_s_bwd_chain_kernel<MMA><<<dim3(2, B), kST, smc, st>>>(ws, sc, B, T, N, C);
            CRW_LAUNCH_RET();
        }
    }
    const size_t smA = (6 * kMat + 128) * sizeof(float);
    if ((rc = set_smem(walk_s_bwd_dA_kernel<MMA>, smA))) return rc;
    walk_s_bwd_dA_kernel<MMA><<<dim3(T - 1, B), kST, smA, st>>>(ws, sc, dA_or_null, B, T, N, C);
    CRW_LAUNCH_RET();
    const size_t smX = (2 * kMat + 2 * kMatX) * sizeof(float);
    if ((rc = set_smem(walk_s_bwd_dx_kernel<MMA>, smX))) return rc;
    walk_s_bwd_dx_kernel<MMA><<<B * T, kST, smX, st>>>(x, ws, sc, dx, B, T, N, C, inv_tau);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// mma = false: fp32 FMA (CRW_PREC_FP32);  mma = true: mma.sync on split bf16 pairs (CRW_PREC_BF16X3 at these sizes)
int walk_small_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st, bool mma) {
    return mma ? walk_small_forward_t<true>(x, B, T, N, C, tau, loss, A_or_null, ws, st)
               : walk_small_forward_t<false>(x, B, T, N, C, tau, loss, A_or_null, ws, st);
}
int walk_small_backward(const float* x, const float* ws, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st, bool mma) {
    return mma ? walk_small_backward_t<true>(x, ws, dloss, dA_or_null, B, T, N, C, tau, dx, sc, st)
               : walk_small_backward_t<false>(x, ws, dloss, dA_or_null, B, T, N, C, tau, dx, sc, st);
}

}  // namespace crw
