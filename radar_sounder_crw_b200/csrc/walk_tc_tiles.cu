// walk_tc_tiles.cu -- tile-parallel tensor-core walk (precision = CRW_PREC_BF16X3), any N.
//
// Same algorithm and reference mapping as walk_f32.cu (src/model.py:22-46 and its autograd).  Here
//   * every GEMM is a grid of independent 128 x 128 output tiles, one CTA each: tcgen05.mma kind::f16 on
//     error-compensated bf16 pairs (x = hi + lo; passes hi.hi, hi.lo, lo.hi), fp32 accumulator in TMEM;
//   * every matrix that is ever a GEMM operand is kept, next to its fp32 copy, as two ROW-MAJOR bf16 planes (hi, lo;
//     row pitch padded to 8 elements, pads zero), written by the epilogue that produces it with 16-byte stores;
//   * an operand used as stored is a K-major tile, an operand used transposed is an MN-major tile of the SAME plane
//     (UMMA majorness bits + LBO/SBO descriptors, pinned by crw_debug_umma_mn_gemm) -- no transposed copies;
//   * staging is TMA: three tensor maps (the E planes, the saved families, the backward families) describe every plane
//     as a stack of matrices, one 64 x 64-element SWIZZLE_128B box is the unit (a K-major tile = two boxes along the rows,
//     an MN-major tile = two boxes along the columns; rows / columns / k past the matrix are zero-filled by the TMA), one
//     elected lane of warp 0 produces three stages ahead (full / empty mbarriers), one elected lane of warp 1 issues the
//     MMAs, the other warps sleep on the accumulator barrier until the epilogue;
//   * S'_t = softmax(A_t^T) is never materialised: the path keeps Q_t = S'_t^T = column-softmax(A_t) (same orientation
//     as A_t, so softmax forward / backward are transposition-free and coalesced) and uses it through the other majorness;
//   * the row-wise work (softmax, cycle cross-entropy, softmax backward, normalise backward) lives in small kernels;
//   * the L / R chains and their adjoints are one launch per step (both chains in one grid): the only serialisation
//     left is the algorithm's own 2(T-3) dependent products forward and backward.
#include "common.cuh"
#include "walk_layout.cuh"
#include "tc_common.cuh"

namespace crw {

constexpr int kTT = 256;
typedef __nv_bfloat16 bf16;

constexpr int kWTile = 128;                       // output tile
constexpr int kWChunk = 64;                       // k elements per stage
constexpr int kWOperand = kWTile * 128;           // one operand plane of one stage: 16 KB (K-major and MN-major alike)
constexpr int kWStage = 4 * kWOperand;            // A_hi, A_lo, B_hi, B_lo
constexpr int kWStages = 3;
constexpr int kWSmem = kWStages * kWStage + 1024;

struct Dims { int B, T, N, C; };

// ---- bf16 arenas ------------------------------------------------------------------------------------------
enum { kFamS = 0, kFamQ, kFamL, kFamR, kFamG, kNumSavedFam };       // saved arena (forward state); Q = S'^T
enum { kFamDL = 0, kFamDR, kFamDA, kNumBwdFam };                    // scratch arena (backward state)

struct TcArena {
    int P;                       // row pitch of N x N planes (N rounded up to 8 elements = 16 bytes)
    int CP;                      // row pitch of the E planes (C rounded up likewise)
    size_t plane;                // B*(T-1)*N*P elements: one plane of one family
    size_t E, fam0, total;       // offsets in bf16 elements
    __host__ __device__ TcArena(const Dims& d, int nfam, bool with_E) {
        P = (d.N + 7) & ~7;
        CP = (d.C + 7) & ~7;
        plane = (size_t)d.B * (d.T - 1) * d.N * P;
        size_t o = 0;
        E = o; if (with_E) o += 2 * (size_t)d.B * d.T * d.N * CP;            // hi, lo   [B*T*N][CP]
        o = (o + 63) & ~size_t(63);
        fam0 = o; o += (size_t)nfam * 2 * plane;                             // hi, lo per family
        total = o;
    }
};
struct Mat2 { bf16 *hi, *lo; int P; };
__device__ __forceinline__ Mat2 mat2(bf16* arena, const TcArena& a, const Dims& d, int fam, int b, int t) {
    bf16* base = arena + a.fam0 + (size_t)fam * 2 * a.plane + ((size_t)b * (d.T - 1) + t) * d.N * a.P;
    return Mat2{base, base + a.plane, a.P};
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void emit_one(const Mat2& mt, int r, int c, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    mt.hi[(size_t)r * mt.P + c] = h;
    mt.lo[(size_t)r * mt.P + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// value for n < N, zero for the pad columns N <= n < pitch
__device__ __forceinline__ void emit_pad(const Mat2& mt, int r, int n, float v, int N) {
    if (n < mt.P) emit_one(mt, r, n, n < N ? v : 0.0f);
}
__device__ __forceinline__ void zero_row_pad(const Mat2& mt, int r, int N, int lane) {
    for (int c = N + lane; c < mt.P; c += 32) { mt.hi[(size_t)r * mt.P + c] = __float2bfloat16_rn(0.f); mt.lo[(size_t)r * mt.P + c] = __float2bfloat16_rn(0.f); }
}

// ---- one 128 x 128 tile from bf16 planes -----------------------------------------------------------------
// K-major use: logical X(r,k) = plane[r*pitch + k]; MN-major use: X(r,k) = plane[k*pitch + r].  An operand is the hi / lo
// matrix pair `zhi`, `zlo` (indices along the third dimension) of one tensor map.
struct OpSrc { const CUtensorMap* map; int zhi, zlo; };
struct TMaps { CUtensorMap E, W, S; };      // E planes [2*B*T][N][C]; saved families [5*2*B*(T-1)][N][N]; backward families [3*2*B*(T-1)][N][N]

struct TcCtx3 { uint8_t* buf; uint64_t* full; uint64_t* empty; uint64_t* acc; uint32_t tmem; uint32_t g, tiles; const TMaps* maps; };

// one operand plane of one stage: 16 KB = two 64 x 64 boxes
template <bool MN>
__device__ __forceinline__ void tma_operand(uint32_t dst, const CUtensorMap* map, int z, int r0, int k0, uint64_t* bar) {
    if (!MN) {
        tc::tma_load_3d(dst, map, k0, r0, z, bar);
        tc::tma_load_3d(dst + 8192, map, k0, r0 + 64, z, bar);
    } else {
        tc::tma_load_3d(dst, map, r0, k0, z, bar);
        tc::tma_load_3d(dst + 8192, map, r0 + 64, k0, z, bar);
    }
}

// epi(m, n, value) is called for every row m < Mvalid of the tile and every column n of the tile (the caller clips n
// against its own extents and pitch).  The accumulator goes TMEM -> registers (thread = row) -> shared memory ->
// registers (warp = row, lane = column) so that every global access of the epilogue is coalesced.
constexpr int kEpPitch = 132;      // floats; 528-byte rows: 16-byte aligned, conflict-free for the v4 stores of phase 1
// acc (TMEM) = (fresh ? 0 : acc) + A B over the whole K extent; several calls may accumulate into one tile.  cx.g counts
// the k-chunks this CTA has ever staged: chunk g lives in stage g % 3, its barriers are in phase (g / 3) & 1.
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void bf_gemm_accumulate(const OpSrc& A, const OpSrc& B, int K, int m0, int n0, TcCtx3& cx, bool fresh) {
    const uint32_t idesc = tc::umma_idesc_bf16_major(kWTile, kWTile, A_MN, B_MN);
    const int nchunks = (K + kWChunk - 1) / kWChunk, warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (tc::elect_one()) {
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t g = cx.g + c, s = g % kWStages;
                if (g >= kWStages) tc::mbar_wait(&cx.empty[s], (g / kWStages - 1) & 1);      // MMAs that read this stage retired
                tc::mbar_arrive_expect_tx(&cx.full[s], kWStage);
                const uint32_t base = tc::smem_u32(cx.buf + s * kWStage);
                tma_operand<A_MN>(base, A.map, A.zhi, m0, c * kWChunk, &cx.full[s]);
                tma_operand<A_MN>(base + kWOperand, A.map, A.zlo, m0, c * kWChunk, &cx.full[s]);
                tma_operand<B_MN>(base + 2 * kWOperand, B.map, B.zhi, n0, c * kWChunk, &cx.full[s]);
                tma_operand<B_MN>(base + 3 * kWOperand, B.map, B.zlo, n0, c * kWChunk, &cx.full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t g = cx.g + c, s = g % kWStages;
                tc::mbar_wait(&cx.full[s], (g / kWStages) & 1);
                tc::tc_fence_after();
                const uint32_t a0 = tc::smem_u32(cx.buf + s * kWStage), b0 = a0 + 2 * kWOperand;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ap = a0 + ((pass == 2) ? kWOperand : 0), bp = b0 + ((pass == 1) ? kWOperand : 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = A_MN ? tc::umma_smem_desc_mn128(ap + ks * 2048, 8192, 1024) : tc::umma_smem_desc_k128(ap + ks * 32);
                        const uint64_t bd = B_MN ? tc::umma_smem_desc_mn128(bp + ks * 2048, 8192, 1024) : tc::umma_smem_desc_k128(bp + ks * 32);
                        tc::umma_bf16_ss(cx.tmem, ad, bd, idesc, (!fresh || (c | pass | ks)) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&cx.empty[s]);
            }
        }
        __syncwarp();
    }
    cx.g += nchunks;
}
template <class Epi>
__device__ __forceinline__ void bf_gemm_epilogue(int m0, int n0, int Mvalid, TcCtx3& cx, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // one commit per tile on a barrier of its own: the warps that took no part in the staging arrive here at once, and a
    // parity wait on a STAGE barrier would alias with that stage's previous use still in flight
    if (warp == 1) {
        if (tc::elect_one()) tc::umma_commit(cx.acc);      // tracks every MMA of this tile: all stages idle when it fires
        __syncwarp();
    }
    tc::mbar_wait(cx.acc, cx.tiles & 1);
    cx.tiles++;
    tc::tc_fence_after();
    float* ep = reinterpret_cast<float*>(cx.buf);
    {
        const int g = warp & 3, half = warp >> 2;
        float* row = ep + (g * 32 + lane) * kEpPitch + half * 64;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tc::tmem_ld_32x32b_x32(cx.tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)(half * 64 + ch * 32), v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(row + ch * 32 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    for (int r = warp; r < kWTile && m0 + r < Mvalid; r += 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) epi(m0 + r, n0 + lane + 32 * q, ep[r * kEpPitch + lane + 32 * q]);
    }
    tc::fence_proxy_async();   // this tile's generic-proxy use of the staging buffer is ordered before the next tile's TMA writes
    __syncthreads();           // the staging buffer and TMEM are reused by the next tile
}
template <bool A_MN, bool B_MN, class Epi>
__device__ __forceinline__ void bf_gemm_tile(const OpSrc& A, const OpSrc& B, int K, int m0, int n0, int Mvalid, TcCtx3& cx, Epi epi) {
    bf_gemm_accumulate<A_MN, B_MN>(A, B, K, m0, n0, cx, true);
    bf_gemm_epilogue(m0, n0, Mvalid, cx, epi);
}

template <class P>
__global__ void __launch_bounds__(kTT, 1) tc_tiles_kernel(const __grid_constant__ P p, const __grid_constant__ TMaps maps) {
    extern __shared__ uint8_t tc_raw[];
    __shared__ uint64_t bars[2 * kWStages + 1];
    __shared__ uint32_t slot;
    TcCtx3 cx;
    cx.buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~uintptr_t(1023));
    cx.full = bars;
    cx.empty = bars + kWStages;
    cx.acc = bars + 2 * kWStages;
    cx.g = 0;
    cx.tiles = 0;
    cx.maps = &maps;
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc<128>(&slot);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * kWStages + 1; ++s) tc::mbar_init(&bars[s], 1);
        tc::fence_barrier_init();
        tc::prefetch_tmap(&maps.E);
        tc::prefetch_tmap(&maps.W);
        tc::prefetch_tmap(&maps.S);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    cx.tmem = slot;
    p.run((int)blockIdx.z, (int)blockIdx.y * kWTile, (int)blockIdx.x * kWTile, cx);
    tc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<128>(cx.tmem);
}

// pointers every problem needs
struct Ctx {
    Dims d;
    float* ws;        // fp32 saved workspace (WalkLayout)
    bf16* wa;         // bf16 saved arena
    float* sc;        // fp32 backward scratch (BwdLayout)
    bf16* sa;         // bf16 backward arena
};
// matrix (fam, b, t) of the saved / backward families, frame (b, t) of the E planes: indices along the maps' third dimension
__device__ __forceinline__ OpSrc op_saved(const TcCtx3& cx, const Dims& d, int fam, int b, int t) {
    const int nm = d.B * (d.T - 1), z = fam * 2 * nm + b * (d.T - 1) + t;
    return OpSrc{&cx.maps->W, z, z + nm};
}
__device__ __forceinline__ OpSrc op_bwd(const TcCtx3& cx, const Dims& d, int fam, int b, int t) {
    const int nm = d.B * (d.T - 1), z = fam * 2 * nm + b * (d.T - 1) + t;
    return OpSrc{&cx.maps->S, z, z + nm};
}
__device__ __forceinline__ OpSrc op_frame(const TcCtx3& cx, const Dims& d, int b, int t) {
    const int z = b * d.T + t;
    return OpSrc{&cx.maps->E, z, z + d.B * d.T};
}

// ---- forward problems -----------------------------------------------------------------------------------
struct AffinityProb {       // batch = b*(T-1) + t :  A_t = E_t E_{t+1}^T / tau
    Ctx c; float* A_out; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int b = z / (d.T - 1), t = z % (d.T - 1), N = d.N;
        const OpSrc A = op_frame(cx, d, b, t), Bm = op_frame(cx, d, b, t + 1);
        float* At = c.ws + lay.mat(lay.A, b, t);
        float* Ao = A_out ? A_out + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
        const float it = inv_tau;
        bf_gemm_tile<false, false>(A, Bm, d.C, m0, n0, N, cx, [&](int m, int n, float v) {
            if (n >= N) return;
            At[(size_t)m * N + n] = v * it;
            if (Ao) Ao[(size_t)m * N + n] = v * it;
        });
    }
};

struct ChainProb {          // batch = role*B + b ; L_k = L_{k-1} S'_{k-1} = L_{k-1} Q_{k-1}^T ;  R_k = S_{k-1} R_{k-1}
    Ctx c; int k;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int role = z / d.B, b = z % d.B, N = d.N;
        if (role == 1 && k < 2) return;
        const OpSrc A = op_saved(cx, d, role == 0 ? kFamL : kFamS, b, k - 1);      // as stored (K-major)
        const OpSrc Bm = op_saved(cx, d, role == 0 ? kFamQ : kFamR, b, k - 1);
        float* out = c.ws + lay.mat(role == 0 ? lay.L : lay.R, b, k);
        const Mat2 om = mat2(c.wa, ar, d, role == 0 ? kFamL : kFamR, b, k);
        auto epi = [&](int m, int n, float v) {
            if (n < N) out[(size_t)m * N + n] = v;
            emit_pad(om, m, n, v, N);
        };
        if (role == 0) bf_gemm_tile<false, false>(A, Bm, N, m0, n0, N, cx, epi);    // B(k,n) = Q[n][k]: K-major
        else bf_gemm_tile<false, true>(A, Bm, N, m0, n0, N, cx, epi);               // B(k,n) = R[k][n]: MN-major
    }
};

struct CycleProb {          // batch = b*K + (k-1) :  M_k = L_k R_k  (raw, into the G slot)
    Ctx c;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const TcArena ar(d, kNumSavedFam, true);
        const int K = d.T - 2, b = z / K, k = z % K + 1, N = d.N;
        float* G = c.ws + lay.mat(lay.G, b, k);
        bf_gemm_tile<false, true>(op_saved(cx, d, kFamL, b, k), op_saved(cx, d, kFamR, b, k), N, m0, n0, N, cx,
                                  [&](int m, int n, float v) { if (n < N) G[(size_t)m * N + n] = v; });
    }
};

// ---- backward problems ----------------------------------------------------------------------------------
// The adjoints of the chain are carried UNSCALED (dL~ = dL / s, dR~ = dR / s with s = dloss / (B N^2)): every product
// below is linear in s, which is applied once in DsProb's epilogue.  That lets one accumulator take both the step's own
// term and the propagated term, and only the bf16 planes of dL~ / dR~ are ever written.
struct BwdChainProb {       // batch = role*B + b ; dL~_j = G_j R_j^T + dL~_{j+1} Q_j ;  dR~_j = L_j^T G_j + S_j^T dR~_{j+1}
    Ctx c; int j;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int role = z / d.B, b = z % d.B, N = d.N, K = d.T - 2;
        if (role == 1 && j < 2) return;
        const OpSrc G = op_saved(cx, d, kFamG, b, j);
        const Mat2 om = mat2(c.sa, ab, d, role == 0 ? kFamDL : kFamDR, b, j);
        if (role == 0) {
            bf_gemm_accumulate<false, false>(G, op_saved(cx, d, kFamR, b, j), N, m0, n0, cx, true);      // B^T(n,k) = R[n][k]
            if (j < K)
                bf_gemm_accumulate<false, true>(op_bwd(cx, d, kFamDL, b, j + 1), op_saved(cx, d, kFamQ, b, j), N, m0, n0, cx, false);
        } else {
            bf_gemm_accumulate<true, true>(op_saved(cx, d, kFamL, b, j), G, N, m0, n0, cx, true);        // A(m,k) = L[k][m]
            if (j < K)
                bf_gemm_accumulate<true, true>(op_saved(cx, d, kFamS, b, j), op_bwd(cx, d, kFamDR, b, j + 1), N, m0, n0, cx, false);
        }
        bf_gemm_epilogue(m0, n0, N, cx, [&](int m, int n, float v) { emit_pad(om, m, n, v, N); });
    }
};

struct DsProb {             // batch = role*B*(T-1) + b*(T-1) + t ; dQ_t = s dL~_{t+1}^T L_t  (= (L_t^T dL_{t+1})^T) ; dS_t = s dR~_{t+1} R_t^T
    Ctx c; const float* dloss;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const WalkLayout lay(d.B, d.T, d.N, d.C);
        const BwdLayout bl(d.B, d.T, d.N);
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int K = d.T - 2, N = d.N, nt = d.T - 1, role = z / (d.B * nt), r = z % (d.B * nt), b = r / nt, t = r % nt;
        const float s = *dloss / ((float)d.B * (float)N * (float)N);
        if (role == 0) {
            if (t + 1 > K) return;
            float* o = c.sc + lay.mat(bl.dSp, b, t);
            bf_gemm_tile<true, true>(op_bwd(cx, d, kFamDL, b, t + 1), op_saved(cx, d, kFamL, b, t), N, m0, n0, N, cx,
                                     [&](int m, int n, float v) { if (n < N) o[(size_t)m * N + n] = v * s; });
        } else {
            if (t < 1 || t + 1 > K) return;
            float* o = c.sc + lay.mat(bl.dS, b, t);
            bf_gemm_tile<false, false>(op_bwd(cx, d, kFamDR, b, t + 1), op_saved(cx, d, kFamR, b, t), N, m0, n0, N, cx,
                                       [&](int m, int n, float v) { if (n < N) o[(size_t)m * N + n] = v * s; });
        }
    }
};

struct DxProb {             // batch = b*T + t ; dE_t = (dA_t E_{t+1} + dA_{t-1}^T E_{t-1}) / tau
    Ctx c; float* dx; float inv_tau;
    __device__ void run(int z, int m0, int n0, TcCtx3& cx) const {
        const Dims& d = c.d;
        const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
        const int b = z / d.T, t = z % d.T, N = d.N, C = d.C;
        float* o = dx + ((size_t)b * d.T + t) * N * C;
        const float it = inv_tau;
        if (t <= d.T - 2) {
            // B(k=j, n=ch) = E_{t+1}[j][ch]: MN-major, MN extent C
            bf_gemm_tile<false, true>(op_bwd(cx, d, kFamDA, b, t), op_frame(cx, d, b, t + 1), N, m0, n0, N, cx,
                                      [&](int m, int n, float v) { if (n < C) o[(size_t)m * C + n] = v * it; });
        } else {
            for (int e = threadIdx.x; e < kWTile * kWTile; e += kTT) {
                const int m = m0 + e / kWTile, ch = n0 + e % kWTile;
                if (m < N && ch < C) o[(size_t)m * C + ch] = 0.0f;
            }
            __syncthreads();
        }
        if (t >= 1) {
            bf_gemm_tile<true, true>(op_bwd(cx, d, kFamDA, b, t - 1), op_frame(cx, d, b, t - 1), N, m0, n0, N, cx,
                                     [&](int m, int n, float v) { if (n < C) o[(size_t)m * C + n] += v * it; });
        }
    }
};

// ---- row-wise kernels -------------------------------------------------------------------------------------
// inverse norms and E = x / ||x|| as bf16 hi/lo [rows][C]
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
    bf16* Elo = Ehi + (size_t)rows * ar.CP;
    for (int ch = lane; ch < ar.CP; ch += 32) {
        const float e = ch < d.C ? xr[ch] * inv : 0.0f;
        const bf16 h = __float2bfloat16_rn(e);
        Ehi[row * ar.CP + ch] = h;
        Elo[row * ar.CP + ch] = __float2bfloat16_rn(e - __bfloat162float(h));
    }
}

__global__ void __launch_bounds__(256) t_identity_kernel(Ctx c) {   // L_0 = I, R_1 = I
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int b = blockIdx.x, N = d.N;
    float* L0 = c.ws + lay.mat(lay.L, b, 0);
    float* R1 = c.ws + lay.mat(lay.R, b, 1);
    const Mat2 l0 = mat2(c.wa, ar, d, kFamL, b, 0), r1 = mat2(c.wa, ar, d, kFamR, b, 1);
    for (size_t i = threadIdx.x; i < (size_t)N * l0.P; i += blockDim.x) {
        const int r = (int)(i / l0.P), cc = (int)(i % l0.P);
        const float v = (r == cc) ? 1.0f : 0.0f;
        if (cc < N) { L0[(size_t)r * N + cc] = v; R1[(size_t)r * N + cc] = v; }
        const bf16 h = __float2bfloat16_rn(v), zz = __float2bfloat16_rn(0.0f);
        l0.hi[i] = h; l0.lo[i] = zz;
        r1.hi[i] = h; r1.lo[i] = zz;
    }
}

// column statistics helper: every warp scans its rows (warp, warp+8, ...) with lanes on consecutive columns (coalesced),
// the 8 partials per column meet in shared memory.  f(i, j) is the value at row i, column j.
template <class F>
__device__ __forceinline__ void column_max(float* out, float* part, int N, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = lane; j < N; j += 32) {
        float m = -INFINITY;
        for (int i = warp; i < N; i += 8) m = fmaxf(m, f(i, j));
        part[warp * N + j] = m;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float m = part[j];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, part[w * N + j]);
        out[j] = m;
    }
    __syncthreads();
}
template <class F>
__device__ __forceinline__ void column_sum(float* out, float* part, int N, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = lane; j < N; j += 32) {
        float a = 0.0f;
        for (int i = warp; i < N; i += 8) a += f(i, j);
        part[warp * N + j] = a;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float a = part[j];
        for (int w = 1; w < 8; ++w) a += part[w * N + j];
        out[j] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) t_softmax_kernel(Ctx c) {    // grid (T-1, B): S_t = rowsoftmax(A_t), Q_t = colsoftmax(A_t)
    extern __shared__ float sm_soft[];      // part[8][N], cmax[N], cinv[N]
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* At = c.ws + lay.mat(lay.A, b, t);
    float* S = c.ws + lay.mat(lay.S, b, t);
    float* Q = c.ws + lay.mat(lay.Sp, b, t);
    const Mat2 ms = mat2(c.wa, ar, d, kFamS, b, t), mq = mat2(c.wa, ar, d, kFamQ, b, t);
    float* part = sm_soft;
    float* cmax = part + 8 * N;
    float* cinv = cmax + N;
    column_max(cmax, part, N, [&](int i, int j) { return At[(size_t)i * N + j]; });
    column_sum(cinv, part, N, [&](int i, int j) { return __expf(At[(size_t)i * N + j] - cmax[j]); });
    for (int j = threadIdx.x; j < N; j += blockDim.x) cinv[j] = 1.0f / cinv[j];
    __syncthreads();
    for (int i = warp; i < N; i += 8) {
        const float* row = At + (size_t)i * N;
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
        mx = warp_max(mx);
        float se = 0.0f;
        for (int j = lane; j < N; j += 32) se += __expf(row[j] - mx);
        se = warp_sum(se);
        const float inv = 1.0f / se;
        for (int j = lane; j < N; j += 32) {
            const float a = row[j];
            const float sv = __expf(a - mx) * inv, qv = __expf(a - cmax[j]) * cinv[j];
            S[(size_t)i * N + j] = sv;
            Q[(size_t)i * N + j] = qv;
            emit_one(ms, i, j, sv);
            emit_one(mq, i, j, qv);
        }
        zero_row_pad(ms, i, N, lane);
        zero_row_pad(mq, i, N, lane);
    }
}

__global__ void __launch_bounds__(256) t_cycle_epi_kernel(Ctx c) {  // grid (T-2, B): G_k = softmax(M_k) - I, loss partial
    __shared__ float red[8];
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const TcArena ar(d, kNumSavedFam, true);
    const int k = blockIdx.x + 1, b = blockIdx.y, N = d.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Gk = c.ws + lay.mat(lay.G, b, k);
    const Mat2 mg = mat2(c.wa, ar, d, kFamG, b, k);
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
            emit_one(mg, r, cc, v);
        }
        zero_row_pad(mg, r, N, lane);
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
    extern __shared__ float sm_da[];   // part[8][N], rS[N], rQ[N]
    const Dims& d = c.d;
    const WalkLayout lay(d.B, d.T, d.N, d.C);
    const BwdLayout bl(d.B, d.T, d.N);
    const TcArena ab(d, kNumBwdFam, false);
    const int t = blockIdx.x, b = blockIdx.y, N = d.N, K = d.T - 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hasQ = (t + 1 <= K), hasS = (t >= 1 && t + 1 <= K);
    const float* S = c.ws + lay.mat(lay.S, b, t);
    const float* Q = c.ws + lay.mat(lay.Sp, b, t);
    const float* dS = c.sc + lay.mat(bl.dS, b, t);
    const float* dQ = c.sc + lay.mat(bl.dSp, b, t);
    const Mat2 ma = mat2(c.sa, ab, d, kFamDA, b, t);
    const float* ext = dA_ext ? dA_ext + ((size_t)b * (d.T - 1) + t) * N * N : nullptr;
    float* part = sm_da;
    float* rS = part + 8 * N;
    float* rQ = rS + N;
    if (hasQ) column_sum(rQ, part, N, [&](int i, int j) { return Q[(size_t)i * N + j] * dQ[(size_t)i * N + j]; });
    for (int i = warp; i < N; i += 8) {
        float a = 0.0f;
        if (hasS) {
            for (int j = lane; j < N; j += 32) a = fmaf(S[(size_t)i * N + j], dS[(size_t)i * N + j], a);
            a = warp_sum(a);
        }
        if (lane == 0) rS[i] = a;
    }
    __syncthreads();
    for (int i = warp; i < N; i += 8) {
        const float ri = rS[i];
        for (int j = lane; j < ma.P; j += 32) {
            float g = 0.0f;
            if (j < N) {
                const size_t e = (size_t)i * N + j;
                g = ext ? ext[e] : 0.0f;
                if (hasS) g += S[e] * (dS[e] - ri);
                if (hasQ) g += Q[e] * (dQ[e] - rQ[j]);
            }
            emit_one(ma, i, j, g);
        }
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
static int launch_tiles(const P& p, const TMaps& maps, int Mrows, int Ncols, int batch, cudaStream_t st) {
    static bool opted[64] = {};   // per instantiation and device; the attribute is a per-device property of the function
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        CRW_CUDA_RET(cudaFuncSetAttribute(tc_tiles_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    if (batch <= 0) return CRW_OK;
    if (batch > 65535) return CRW_ERR_UNSUPPORTED;       // the batch index rides in grid.z (the fp32 engine has no such limit)
    dim3 grid(ceil_div(Ncols, kWTile), ceil_div(Mrows, kWTile), batch);
    tc_tiles_kernel<P><<<grid, kTT, kWSmem, st>>>(p, maps);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// the column-statistics kernels keep 10 N floats in shared memory
template <class Kern>
static int opt_in_rowwise_smem(Kern kern, int N) {
    const size_t need = 10 * (size_t)N * sizeof(float);
    if (need > 227 * 1024) return CRW_ERR_UNSUPPORTED;
    if (need > 48 * 1024) CRW_CUDA_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    return CRW_OK;
}

// bf16 operand planes kept next to the fp32 state
size_t walk_tiles_saved_extra_bytes(int B, int T, int N, int C) {
    return TcArena(Dims{B, T, N, C}, kNumSavedFam, true).total * sizeof(bf16) + 256;
}
size_t walk_tiles_scratch_extra_bytes(int B, int T, int N, int C) {
    return TcArena(Dims{B, T, N, C}, kNumBwdFam, false).total * sizeof(bf16) + 256;
}
static bf16* arena_after(float* f32_base, size_t f32_floats) {
    return reinterpret_cast<bf16*>((reinterpret_cast<uintptr_t>(f32_base + f32_floats) + 255) & ~uintptr_t(255));
}
// tensor maps over the bf16 arenas (sa = nullptr: forward, no backward families yet)
static int make_tile_maps(TMaps* m, const Dims& d, bf16* wa, bf16* sa) {
    const TcArena ar(d, kNumSavedFam, true), ab(d, kNumBwdFam, false);
    const uint64_t nm = (uint64_t)d.B * (d.T - 1);
    memset(m, 0, sizeof(*m));
    int rc = make_tmap_bf16_mats(&m->E, wa + ar.E, d.C, d.N, 2 * (uint64_t)d.B * d.T, ar.CP);
    if (rc == CRW_OK) rc = make_tmap_bf16_mats(&m->W, wa + ar.fam0, d.N, d.N, kNumSavedFam * 2 * nm, ar.P);
    if (rc == CRW_OK) {
        if (sa) rc = make_tmap_bf16_mats(&m->S, sa + ab.fam0, d.N, d.N, kNumBwdFam * 2 * nm, ab.P);
        else m->S = m->W;
    }
    return rc;
}

int walk_tiles_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, float* ws,
                       cudaStream_t st) {
    const Dims d{B, T, N, C};
    const WalkLayout lay(B, T, N, C);
    Ctx c{d, ws, arena_after(ws, lay.total), nullptr, nullptr};
    int rc;
    TMaps maps;
    if ((rc = make_tile_maps(&maps, d, c.wa, nullptr))) return rc;
    const long long rows = (long long)B * T * N;
    t_rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, c);
    CRW_LAUNCH_RET();
    if ((rc = launch_tiles(AffinityProb{c, A_or_null, 1.0f / tau}, maps, N, N, B * (T - 1), st))) return rc;
    if (T < 3) {
        t_zero_loss_kernel<<<1, 1, 0, st>>>(loss);
        CRW_LAUNCH_RET();
        return CRW_OK;
    }
    if ((rc = opt_in_rowwise_smem(t_softmax_kernel, N))) return rc;
    t_softmax_kernel<<<dim3(T - 1, B), 256, 10 * N * sizeof(float), st>>>(c);
    CRW_LAUNCH_RET();
    t_identity_kernel<<<B, 256, 0, st>>>(c);
    CRW_LAUNCH_RET();
    const int K = T - 2;
    for (int k = 1; k <= K; ++k)
        if ((rc = launch_tiles(ChainProb{c, k}, maps, N, N, k >= 2 ? 2 * B : B, st))) return rc;
    if ((rc = launch_tiles(CycleProb{c}, maps, N, N, B * K, st))) return rc;
    t_cycle_epi_kernel<<<dim3(K, B), 256, 0, st>>>(c);
    CRW_LAUNCH_RET();
    t_loss_reduce_kernel<<<1, 32, 0, st>>>(ws, loss, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

int walk_tiles_backward(const float* x, const float* ws_c, const float* dloss, const float* dA_or_null, int B, int T, int N, int C,
                        float tau, float* dx, float* sc, cudaStream_t st) {
    const Dims d{B, T, N, C};
    const WalkLayout lay(B, T, N, C);
    const BwdLayout bl(B, T, N);
    float* ws = const_cast<float*>(ws_c);
    Ctx c{d, ws, arena_after(ws, lay.total), sc, arena_after(sc, bl.total)};
    int rc;
    TMaps maps;
    if ((rc = make_tile_maps(&maps, d, c.wa, c.sa))) return rc;
    const int K = T - 2;
    if (T >= 3) {
        for (int j = K; j >= 1; --j)
            if ((rc = launch_tiles(BwdChainProb{c, j}, maps, N, N, j >= 2 ? 2 * B : B, st))) return rc;
        if ((rc = launch_tiles(DsProb{c, dloss}, maps, N, N, 2 * B * (T - 1), st))) return rc;
    }
    if ((rc = opt_in_rowwise_smem(t_dA_epi_kernel, N))) return rc;
    t_dA_epi_kernel<<<dim3(T - 1, B), 256, 10 * N * sizeof(float), st>>>(c, dA_or_null);
    CRW_LAUNCH_RET();
    if ((rc = launch_tiles(DxProb{c, dx, 1.0f / tau}, maps, N, C, B * T, st))) return rc;
    t_dx_epi_kernel<<<dim3(T, B), 256, 0, st>>>(x, ws, dx, d);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

}  // namespace crw
