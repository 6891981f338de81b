// walk_tc_tiles.cu -- tile-parallel tensor-core walk (precision = CRW_PREC_BF16X3), any N.
//
// Same algorithm and reference mapping as walk_f32.cu (src/model.py:22-46 and its autograd).  Here
//   * every GEMM is a grid of independent 128 x 128 output tiles, one CTA each: tcgen05.mma kind::f16 on
//     error-compensated bf16 pairs (x = hi + lo; passes hi.hi, hi.lo, lo.hi), fp32 accumulator in TMEM;
//   * every matrix that is ever a GEMM operand is kept, next to its fp32 copy, as bf16 hi/lo planes in BOTH
//     orientations (X and X^T, row pitch padded to 8), written by the epilogue that produces it, so that an operand
//     tile is always "128 rows x 64 contiguous k": staging is sixteen 16-byte cp.async per thread per k-chunk straight
//     into the UMMA SWIZZLE_128B layout, double buffered against the MMAs -- no conversion, no transposition, no
//     scalar loads on the GEMM path;
//   * the row-wise work (softmax, cycle cross-entropy, softmax backward, normalise backward) lives in small kernels;
//   * the L / R chains and their adjoints are one launch per step (both chains in one grid): the only serialisation
//     left is the algorithm's own 2(T-3) dependent products forward and backward.
#include "common.cuh"
#include "walk_layout.cuh"
#include "walk_tc.cuh"

namespace crw {

constexpr int kTT = 256;
typedef __nv_bfloat16 bf16;

struct Dims { int B, T, N, C; };

// ---- bf16 arenas ------------------------------------------------------------------------------------------
enum { kFamS = 0, kFamSp, kFamL, kFamR, kFamG, kNumSavedFam };      // saved arena (forward state)
enum { kFamDL = 0, kFamDR, kFamDA, kNumBwdFam };                    // scratch arena (backward state)

struct TcArena {
    int P;                       // row pitch of N x N planes (N rounded up to 8 elements = 16 bytes)
    size_t plane;                // B*(T-1)*N*P elements: one plane of one family
    size_t E, ET, fam0, total;   // offsets in bf16 elements
    __host__ __device__ TcArena(const Dims& d, int nfam, bool with_E) {
        P = (d.N + 7) & ~7;
        plane = (size_t)d.B * (d.T - 1) * d.N * P;
        size_t o = 0;
        E = o;  if (with_E) o += 2 * (size_t)d.B * d.T * d.N * d.C;          // hi, lo        [B*T*N][C]
        ET = o; if (with_E) o += 2 * (size_t)d.B * d.T * d.C * P;            // hi, lo        [B*T][C][P]
        o = (o + 63) & ~size_t(63);
        fam0 = o; o += (size_t)nfam * 4 * plane;                             // hi, lo, hiT, loT per family
        total = o;
    }
};
struct Mat4 { bf16 *hi, *lo, *hiT, *loT; int P; };
__device__ __forceinline__ Mat4 mat4(bf16* arena, const TcArena& a, const Dims& d, int fam, int b, int t) {
    bf16* base = arena + a.fam0 + (size_t)fam * 4 * a.plane + ((size_t)b * (d.T - 1) + t) * d.N * a.P;
    return Mat4{base, base + a.plane, base + 2 * a.plane, base + 3 * a.plane, a.P};
}
__device__ __forceinline__ void emit(const Mat4& m, int r, int c, float v) {   // X[r][c] and X^T[c][r]
    const bf16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    m.hi[(size_t)r * m.P + c] = h;   m.lo[(size_t)r * m.P + c] = l;
    m.hiT[(size_t)c * m.P + r] = h;  m.loT[(size_t)c * m.P + r] = l;
}

// ---- one 128 x 128 tile from bf16 K-major sources ------------------------------------------------------------
struct OpSrc { const bf16 *hi, *lo; int pitch, rows; };     // logical [rows][K], k contiguous, pad columns are zero

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void stage_chunk_async(uint8_t* st, const OpSrc& A, const OpSrc& B, int m0, int n0, int k0, int K) {
    const uint32_t base = tc::smem_u32(st);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int id = threadIdx.x + it * kTT;          // 1024 chunks of 8 elements per operand plane
        const int r = id >> 3, c = id & 7, kk = k0 + c * 8;
        const uint32_t off = (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ (r & 7)) & 7) << 4));
        const bool ka = kk < K;
        const bool va = ka && (m0 + r) < A.rows, vb = ka && (n0 + r) < B.rows;
        const size_t ao = va ? (size_t)(m0 + r) * A.pitch + kk : 0, bo = vb ? (size_t)(n0 + r) * B.pitch + kk : 0;
        cp_async16_zfill(base + off, A.hi + ao, va);
        cp_async16_zfill(base + kTcOperandBytes + off, A.lo + ao, va);
        cp_async16_zfill(base + 2 * kTcOperandBytes + off, B.hi + bo, vb);
        cp_async16_zfill(base + 3 * kTcOperandBytes + off, B.lo + bo, vb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <class Epi>
__device__ __forceinline__ void bf_gemm_tile(const OpSrc& A, const OpSrc& B, int K, int m0, int n0, TcGemmCtx& cx, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idesc = tc::umma_idesc_bf16(kTcTile, kTcTile);
    const int nchunks = (K + kTcKChunk - 1) / kTcKChunk;
    if (cx.uses[0] > 0) tc::mbar_wait(&cx.bar[0], (cx.uses[0] - 1) & 1);
    stage_chunk_async(cx.buf, A, B, m0, n0, 0, K);
    int last_stage = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        if (c + 1 < nchunks) {
            const int s2 = s ^ 1;
            if (cx.uses[s2] > 0) tc::mbar_wait(&cx.bar[s2], (cx.uses[s2] - 1) & 1);   // MMAs that read this stage retired
            stage_chunk_async(cx.buf + s2 * kTcStageBytes, A, B, m0, n0, (c + 1) * kTcKChunk, K);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        tc::fence_proxy_async();       // cp.async (generic proxy) writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::tc_fence_after();
            const uint32_t a0 = tc::smem_u32(cx.buf + s * kTcStageBytes), b0 = a0 + 2 * kTcOperandBytes;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ap = a0 + ((pass == 2) ? kTcOperandBytes : 0), bp = b0 + ((pass == 1) ? kTcOperandBytes : 0);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    tc::umma_bf16_ss(cx.tmem, tc::umma_smem_desc_k128(ap + ks * 32), tc::umma_smem_desc_k128(bp + ks * 32),
                                     idesc, (c | pass | ks) ? 1u : 0u);
            }
            tc::umma_commit(&cx.bar[s]);
        }
        cx.uses[s]++;
        last_stage = s;
    }
    tc::mbar_wait(&cx.bar[last_stage], (cx.uses[last_stage] - 1) & 1);
    tc::tc_fence_after();
    const int g = warp & 3, half = warp >> 2;
    const int m = m0 + g * 32 + lane;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        const int cb = half * 64 + ch * 32;
        tc::tmem_ld_32x32b_x32(cx.tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)cb, v);
        tc::tmem_ld_wait();
        if (m < A.rows) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = n0 + cb + i;
                if (n < B.rows) epi(m, n, v[i]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
}

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

// pointers every problem needs
struct Ctx {
    Dims d;
    float* ws;        // fp32 saved workspace (WalkLayout)
    bf16* wa;         // bf16 saved arena
    float* sc;        // fp32 backward scratch (BwdLayout)
    bf16* sa;         // bf16 backward arena
};
__device__ __forceinline__ OpSrc src_X(const Mat4& m, int rows) { return OpSrc{m.hi, m.lo, m.P, rows}; }
__device__ __forceinline__ OpSrc src_XT(const Mat4& m, int rows) { return OpSrc{m.hiT, m.loT, m.P, rows}; }

// ---- forward problems -----------------------------------------------------------------------------------
struct AffinityProb {       // batch = b*(T-1) + t :  A_t = E_t E_{t+1}^T / tau
    Ctx c; float* A_out; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int b = z / (d.T - 1), t = z % (d.T - 1), N = d.N;
        const size_t rows = (size_t)d.B * d.T * N, r0 = ((size_t)b * d.T + t) * N;
        const bf16* Ehi = c.wa + ar.E;
        const bf16* Elo = Ehi + rows * d.C;
        const OpSrc A{Ehi + r0 * d.C, Elo + r0 * d.C, d.C, N}, Bm{Ehi + (r0 + N) * d.C, Elo + (r0 + N) * d.C, d.C, N};
        float* At = c.ws + lay.mat(lay.A, b, t);
        float* Ao = A_out ? A_out + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
        const float it = inv_tau;
        bf_gemm_tile(A, Bm, d.C, m0, n0, cx, [&](int m, int n, float v) {
            const float a = v * it;
            At[(size_t)m * N + n] = a;
            if (Ao) Ao[(size_t)m * N + n] = a;
        });
    }
};

struct ChainProb {          // batch = role*B + b ; L_k = L_{k-1} S'_{k-1} ;  R_k = S_{k-1} R_{k-1}
    Ctx c; int k;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int role = z / d.B, b = z % d.B, N = d.N;
        if (role == 1 && k < 2) return;
        const OpSrc A = role == 0 ? src_X(mat4(c.wa, ar, d, kFamL, b, k - 1), N) : src_X(mat4(c.wa, ar, d, kFamS, b, k - 1), N);
        const OpSrc Bm = role == 0 ? src_XT(mat4(c.wa, ar, d, kFamSp, b, k - 1), N) : src_XT(mat4(c.wa, ar, d, kFamR, b, k - 1), N);
        float* out = c.ws + lay.mat(role == 0 ? lay.L : lay.R, b, k);
        const Mat4 om = mat4(c.wa, ar, d, role == 0 ? kFamL : kFamR, b, k);
        bf_gemm_tile(A, Bm, N, m0, n0, cx, [&](int m, int n, float v) { out[(size_t)m * N + n] = v; emit(om, m, n, v); });
    }
};

struct CycleProb {          // batch = b*K + (k-1) :  M_k = L_k R_k  (raw, into the G slot)
    Ctx c;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int K = d.T - 2, b = z / K, k = z % K + 1, N = d.N;
        float* G = c.ws + lay.mat(lay.G, b, k);
        bf_gemm_tile(src_X(mat4(c.wa, ar, d, kFamL, b, k), N), src_XT(mat4(c.wa, ar, d, kFamR, b, k), N), N, m0, n0, cx,
                     [&](int m, int n, float v) { G[(size_t)m * N + n] = v; });
    }
};

// ---- backward problems ----------------------------------------------------------------------------------
struct OwnProb {            // batch = role*B*K + b*K + (k-1) :  dL_k = s G_k R_k^T ;  dR_k = s L_k^T G_k
    Ctx c; const float* dloss;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int K = d.T - 2, N = d.N, role = z / (d.B * K), r = z % (d.B * K), b = r / K, k = r % K + 1;
        const float s = *dloss / ((float)d.B * (float)N * (float)N);
        const Mat4 G = mat4(c.wa, ar, d, kFamG, b, k);
        if (role == 0) {
            float* o = c.sc + lay.mat(bl.dL, b, k);
            const Mat4 om = mat4(c.sa, ab, d, kFamDL, b, k);
            bf_gemm_tile(src_X(G, N), src_X(mat4(c.wa, ar, d, kFamR, b, k), N), N, m0, n0, cx, [&](int m, int n, float v) {
                const float w = v * s;
                o[(size_t)m * N + n] = w;
                emit(om, m, n, w);
            });
        } else {
            float* o = c.sc + lay.mat(bl.dR, b, k);
            const Mat4 om = mat4(c.sa, ab, d, kFamDR, b, k);
            bf_gemm_tile(src_XT(mat4(c.wa, ar, d, kFamL, b, k), N), src_XT(G, N), N, m0, n0, cx, [&](int m, int n, float v) {
                const float w = v * s;
                o[(size_t)m * N + n] = w;
                emit(om, m, n, w);
            });
        }
    }
};

struct BwdChainProb {       // batch = role*B + b ; dL_j += dL_{j+1} S'_j^T ; dR_j += S_j^T dR_{j+1}
    Ctx c; int j;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int role = z / d.B, b = z % d.B, N = d.N;
        if (role == 1 && j < 2) return;
        if (role == 0) {
            float* o = c.sc + lay.mat(bl.dL, b, j);
            const Mat4 om = mat4(c.sa, ab, d, kFamDL, b, j);
            bf_gemm_tile(src_X(mat4(c.sa, ab, d, kFamDL, b, j + 1), N), src_X(mat4(c.wa, ar, d, kFamSp, b, j), N), N, m0, n0, cx,
                         [&](int m, int n, float v) {
                             const float w = o[(size_t)m * N + n] + v;
                             o[(size_t)m * N + n] = w;
                             emit(om, m, n, w);
                         });
        } else {
            float* o = c.sc + lay.mat(bl.dR, b, j);
            const Mat4 om = mat4(c.sa, ab, d, kFamDR, b, j);
            bf_gemm_tile(src_XT(mat4(c.wa, ar, d, kFamS, b, j), N), src_XT(mat4(c.sa, ab, d, kFamDR, b, j + 1), N), N, m0, n0, cx,
                         [&](int m, int n, float v) {
                             const float w = o[(size_t)m * N + n] + v;
                             o[(size_t)m * N + n] = w;
                             emit(om, m, n, w);
                         });
        }
    }
};

struct DsProb {             // batch = role*B*(T-1) + b*(T-1) + t ; dS'_t = L_t^T dL_{t+1} ; dS_t = dR_{t+1} R_t^T
    Ctx c;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int K = d.T - 2, N = d.N, nt = d.T - 1, role = z / (d.B * nt), r = z % (d.B * nt), b = r / nt, t = r % nt;
        if (role == 0) {
            if (t + 1 > K) return;
            float* o = c.sc + lay.mat(bl.dSp, b, t);
            bf_gemm_tile(src_XT(mat4(c.wa, ar, d, kFamL, b, t), N), src_XT(mat4(c.sa, ab, d, kFamDL, b, t + 1), N), N, m0, n0, cx,
                         [&](int m, int n, float v) { o[(size_t)m * N + n] = v; });
        } else {
            if (t < 1 || t + 1 > K) return;
            float* o = c.sc + lay.mat(bl.dS, b, t);
            bf_gemm_tile(src_X(mat4(c.sa, ab, d, kFamDR, b, t + 1), N), src_X(mat4(c.wa, ar, d, kFamR, b, t), N), N, m0, n0, cx,
                         [&](int m, int n, float v) { o[(size_t)m * N + n] = v; });
        }
    }
};

struct DxProb {             // batch = b*T + t ; dE_t = (dA_t E_{t+1} + dA_{t-1}^T E_{t-1}) / tau
    Ctx c; float* dx; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcGemmCtx& cx) const {
        const Dims& d = c.d;
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int b = z / d.T, t = z % d.T, N = d.N, C = d.C;
        float* o = dx + ((size_t)b * d.T + t) * N * C;
        const bf16* EThi = c.wa + ar.ET;
        const bf16* ETlo = EThi + (size_t)d.B * d.T * C * ar.P;
        const float it = inv_tau;
        if (t <= d.T - 2) {
            const size_t eo = ((size_t)b * d.T + t + 1) * C * ar.P;
            bf_gemm_tile(src_X(mat4(c.sa, ab, d, kFamDA, b, t), N), OpSrc{EThi + eo, ETlo + eo, ar.P, C}, N, m0, n0, cx,
                         [&](int m, int ch, float v) { o[(size_t)m * C + ch] = v * it; });
        } else {
            for (int e = threadIdx.x; e < kTcTile * kTcTile; e += kTT) {
                const int m = m0 + e / kTcTile, ch = n0 + e % kTcTile;
                if (m < N && ch < C) o[(size_t)m * C + ch] = 0.0f;
            }
            __syncthreads();
        }
        if (t >= 1) {
            const size_t eo = ((size_t)b * d.T + t - 1) * C * ar.P;
            bf_gemm_tile(src_XT(mat4(c.sa, ab, d, kFamDA, b, t - 1), N), OpSrc{EThi + eo, ETlo + eo, ar.P, C}, N, m0, n0, cx,
                         [&](int m, int ch, float v) { o[(size_t)m * C + ch] += v * it; });
        }
    }
};

// ---- row-wise kernels -------------------------------------------------------------------------------------
// pad columns [N, P) of every bf16 plane must be zero (they are fetched by the 16-byte operand chunks)
__global__ void __launch_bounds__(256) t_zero_pads_kernel(bf16* planes, size_t n_rows, int N, int P) {
    const int pad = P - N;
    const size_t total = n_rows * pad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        planes[(i / pad) * P + N + (i % pad)] = __float2bfloat16_rn(0.0f);
}

// inverse norms, E = x / ||x|| as bf16 hi/lo [rows][C], and E^T hi/lo [b,t][C][P]
__global__ void __launch_bounds__(256) t_rownorm_kernel(const float* __restrict__ x, Ctx c) {
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), rows = (long long)d.B * d.T * d.N;
    if (row >= rows) return;
    const float* xr = x + row * d.C;
    float ss = 0.0f;
    for (int ch = lane; ch < d.C; ch += 32) ss = fmaf(xr[ch], xr[ch], ss);
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), kNormEps);
    if (lane == 0) c.ws[lay.invn + row] = inv;
    bf16* Ehi = c.wa + ar.E;
    bf16* Elo = Ehi + (size_t)rows * d.C;
    bf16* EThi = c.wa + ar.ET;
    bf16* ETlo = EThi + (size_t)d.B * d.T * d.C * ar.P;
    const long long bt = row / d.N;
    const int n = (int)(row % d.N);
    for (int ch = lane; ch < d.C; ch += 32) {
        const float e = xr[ch] * inv;
        const bf16 h = __float2bfloat16_rn(e), l = __float2bfloat16_rn(e - __bfloat162float(h));
        Ehi[row * d.C + ch] = h;
        Elo[row * d.C + ch] = l;
        EThi[((size_t)bt * d.C + ch) * ar.P + n] = h;
        ETlo[((size_t)bt * d.C + ch) * ar.P + n] = l;
    }
}

__global__ void __launch_bounds__(256) t_identity_kernel(Ctx c) {   // L_0 = I, R_1 = I
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int b = blockIdx.x, N = d.N;
    float* L0 = c.ws + lay.mat(lay.L, b, 0);
    float* R1 = c.ws + lay.mat(lay.R, b, 1);
    const Mat4 l0 = mat4(c.wa, ar, d, kFamL, b, 0), r1 = mat4(c.wa, ar, d, kFamR, b, 1);
    for (size_t i = threadIdx.x; i < (size_t)N * N; i += blockDim.x) {
        const int r = (int)(i / N), cc = (int)(i % N);
        const float v = (r == cc) ? 1.0f : 0.0f;
        L0[i] = v;
        R1[i] = v;
        const bf16 h = __float2bfloat16_rn(v), z = __float2bfloat16_rn(0.0f);
        l0.hi[(size_t)r * l0.P + cc] = h; l0.lo[(size_t)r * l0.P + cc] = z;     // only the X plane of L_0 ...
        r1.hiT[(size_t)r * r1.P + cc] = h; r1.loT[(size_t)r * r1.P + cc] = z;   // ... and the X^T plane of R_1 are read
        r1.hi[(size_t)r * r1.P + cc] = h; r1.lo[(size_t)r * r1.P + cc] = z;
        l0.hiT[(size_t)r * l0.P + cc] = h; l0.loT[(size_t)r * l0.P + cc] = z;
    }
}

__global__ void __launch_bounds__(256) t_softmax_kernel(Ctx c) {    // grid (T-1, B): S_t, S'_t from A_t
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* At = c.ws + lay.mat(lay.A, b, t);
    float* S = c.ws + lay.mat(lay.S, b, t);
    float* Sp = c.ws + lay.mat(lay.Sp, b, t);
    const Mat4 ms = mat4(c.wa, ar, d, kFamS, b, t), mp = mat4(c.wa, ar, d, kFamSp, b, t);
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
        const Mat4& mm = col ? mp : ms;
        for (int j = lane; j < N; j += 32) {
            const float v = __expf(At[base + j * step] - mx) * inv;
            dst[j] = v;
            emit(mm, i, j, v);
        }
    }
}

__global__ void __launch_bounds__(256) t_cycle_epi_kernel(Ctx c) {  // grid (T-2, B): G_k = softmax(M_k) - I, loss partial
    __shared__ float red[8];
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int k = blockIdx.x + 1, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Gk = c.ws + lay.mat(lay.G, b, k);
    const Mat4 mg = mat4(c.wa, ar, d, kFamG, b, k);
    float part = 0.0f;
    for (int r = warp; r < N; r += 8) {
        float* row = Gk + (size_t)r * N;
        float mx = -INFINITY;
        for (int cc = lane; cc < N; cc += 32) mx = fmaxf(mx, row[cc]);
        mx = warp_max(mx);
        float se = 0.0f;
        for (int cc = lane; cc < N; cc += 32) se += __expf(row[cc] - mx);
        se = warp_sum(se);
        const float diag = row[r];
        __syncwarp();
        const float inv = 1.0f / se;
        for (int cc = lane; cc < N; cc += 32) {
            const float v = __expf(row[cc] - mx) * inv - (cc == r ? 1.0f : 0.0f);
            row[cc] = v;
            emit(mg, r, cc, v);
        }
        part += (logf(se) + mx) - diag;
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w];
        c.ws[lay.part + (size_t)b * (d.T - 1) + k] = s;
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

__global__ void __launch_bounds__(256) t_dA_epi_kernel(Ctx c, const float* dA_ext) {   // grid (T-1, B)
    extern __shared__ float rdot[];   // rS[N], rSp[N]
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const BwdLayout bl(d.B, d.T, d.N);
    const TcArena ab(d, kNumBwdFam, false);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, K = d.T - 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hasSp = (t + 1 <= K), hasS = (t >= 1 && t + 1 <= K);
    const float* S = c.ws + lay.mat(lay.S, b, t);
    const float* Sp = c.ws + lay.mat(lay.Sp, b, t);
    const float* dS = c.sc + lay.mat(bl.dS, b, t);
    const float* dSp = c.sc + lay.mat(bl.dSp, b, t);
    const Mat4 ma = mat4(c.sa, ab, d, kFamDA, b, t);
    const float* ext = dA_ext ? dA_ext + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
    for (int r = warp; r < 2 * N; r += 8) {
        const bool second = r >= N;
        const int i = second ? r - N : r;
        float a = 0.0f;
        if (second ? hasSp : hasS) {
            const float* Pm = (second ? Sp : S) + (size_t)i * N;
            const float* dP = (second ? dSp : dS) + (size_t)i * N;
            for (int j = lane; j < N; j += 32) a = fmaf(Pm[j], dP[j], a);
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
        emit(ma, i, j, g);
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
            for (int ch = lane; ch < C; ch += 32) orow[ch] *= inv;
            continue;
        }
        float dot = 0.0f;
        for (int ch = lane; ch < C; ch += 32) dot = fmaf(xr[ch] * inv, orow[ch], dot);
        dot = warp_sum(dot);
        for (int ch = lane; ch < C; ch += 32) orow[ch] = (orow[ch] - xr[ch] * inv * dot) * inv;
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

size_t walk_tiles_saved_extra_bytes(int B, int T, int N, int C) {
    return TcArena(Dims{B, T, N, C}, kNumSavedFam, true).total * sizeof(bf16) + 256;
}
size_t walk_tiles_scratch_extra_bytes(int B, int T, int N, int C) {
    return TcArena(Dims{B, T, N, C}, kNumBwdFam, false).total * sizeof(bf16) + 256;
}
static bf16* arena_after(float* f32_base, size_t f32_floats) {
    return reinterpret_cast<bf16*>((reinterpret_cast<uintptr_t>(f32_base + f32_floats) + 255) & ~uintptr_t(255));
}
static int zero_pads(bf16* arena, const TcArena& a, const Dims& d, int nfam, bool with_ET, cudaStream_t st) {
    if (a.P == d.N) return CRW_OK;
    const size_t fam_rows = (size_t)nfam * 4 * d.B * (d.T - 1) * d.N;
    t_zero_pads_kernel<<<148 * 4, 256, 0, st>>>(arena + a.fam0, fam_rows, d.N, a.P);
    CRW_LAUNCH_RET();
    if (with_ET) {
        t_zero_pads_kernel<<<148 * 2, 256, 0, st>>>(arena + a.ET, (size_t)2 * d.B * d.T * d.C, d.N, a.P);
        CRW_LAUNCH_RET();
    }
    return CRW_OK;
}

int walk_tiles_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st) {
    if (C % 8) return CRW_ERR_ALIGN;      // 16-byte operand chunks along the channel axis
    const Dims d{B, T, N, C};
    const WalkLayout lay(B, T, N, C);
    const TcArena ar(d, kNumSavedFam, true);
    Ctx c{d, ws, arena_after(ws, lay.total), nullptr, nullptr};
    int rc = zero_pads(c.wa, ar, d, kNumSavedFam, true, st);
    if (rc) return rc;
    const long long rows = (long long)B * T * N;
    t_rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, c);
    CRW_LAUNCH_RET();
    if ((rc = launch_tiles(AffinityProb{c, A_or_null, 1.0f / tau}, N, N, B * (T - 1), st))) return rc;
    if (T < 3) {
        t_zero_loss_kernel<<<1, 1, 0, st>>>(loss);
        CRW_LAUNCH_RET();
        return CRW_OK;
    }
    t_softmax_kernel<<<dim3(T - 1, B), 256, 0, st>>>(c);
    CRW_LAUNCH_RET();
    t_identity_kernel<<<B, 256, 0, st>>>(c);
    CRW_LAUNCH_RET();
    const int K = T - 2;
    for (int k = 1; k <= K; ++k)
        if ((rc = launch_tiles(ChainProb{c, k}, N, N, k >= 2 ? 2 * B : B, st))) return rc;
    if ((rc = launch_tiles(CycleProb{c}, N, N, B * K, st))) return rc;
    t_cycle_epi_kernel<<<dim3(K, B), 256, 0, st>>>(c);
    CRW_LAUNCH_RET();
    t_loss_reduce_kernel<<<1, 32, 0, st>>>(ws, loss, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

int walk_tiles_backward(const float* x, const float* ws_c, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st) {
    if (C % 8) return CRW_ERR_ALIGN;
    const Dims d{B, T, N, C};
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    const TcArena ab(d, kNumBwdFam, false);
    float* ws = const_cast<float*>(ws_c);
    Ctx c{d, ws, arena_after(ws, lay.total), sc, arena_after(sc, bl.total)};
    int rc = zero_pads(c.sa, ab, d, kNumBwdFam, false, st);
    if (rc) return rc;
    const int K = T - 2;
    if (T >= 3) {
        if ((rc = launch_tiles(OwnProb{c, dloss}, N, N, 2 * B * K, st))) return rc;
        for (int j = K - 1; j >= 1; --j)
            if ((rc = launch_tiles(BwdChainProb{c, j}, N, N, j >= 2 ? 2 * B : B, st))) return rc;
        if ((rc = launch_tiles(DsProb{c}, N, N, 2 * B * (T - 1), st))) return rc;
    }
    t_dA_epi_kernel<<<dim3(T - 1, B), 256, 2 * N * sizeof(float), st>>>(c, dA_or_null);
    CRW_LAUNCH_RET();
    if ((rc = launch_tiles(DxProb{c, dx, 1.0f / tau}, N, C, B * T, st))) return rc;
    t_dx_epi_kernel<<<dim3(T, B), 256, 0, st>>>(x, ws, dx, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

}  // namespace crw
