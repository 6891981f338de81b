// walk_tc.cuh -- CTA-wide tensor-core GEMM tile for the walk (any N): tcgen05.mma, TMEM accumulator.
//
// C[128 x 128 tile] = op(A) * op(B) with fp32 operands in global memory.  The operands of the walk are produced in
// fp32 by the previous stage's epilogue (softmax, chain products, adjoints), some of them consumed transposed, so
// instead of TMA the CTA stages them itself: coalesced loads -> error-compensated bf16 split (x = hi + lo) ->
// 16-byte chunks written straight into the UMMA K-major SWIZZLE_128B layout (chunk index xor row%8), double
// buffered so that staging chunk c+1 overlaps the 12 MMAs (3 passes hi.hi, hi.lo, lo.hi x 4 k-steps) of chunk c.
// The accumulator stays in TMEM across the whole K loop and is read back once with tcgen05.ld for the epilogue
// (scale / accumulate).  Used by walk_tc_tiles.cu, one CTA per 128 x 128 output tile.
#pragma once
#include "tc_common.cuh"

namespace crw {

constexpr int kTcTile = 128;                 // output tile (M = N = 128)
constexpr int kTcKChunk = 64;                // bf16 elements per 128-byte swizzled row
constexpr int kTcOperandBytes = kTcTile * 128;            // one [128 rows][64 bf16] sub-tile = 16 KB
constexpr int kTcStageBytes = 4 * kTcOperandBytes;        // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr int kTcSmemBytes = 2 * kTcStageBytes + 1024;    // two stages + alignment slack

struct TcGemmCtx {
    uint8_t* buf;          // 1024-byte aligned, 2 stages
    uint64_t* bar;         // [2] "MMAs reading this stage have retired"
    uint32_t tmem;         // TMEM base (128 fp32 columns)
    uint32_t uses[2];      // completed-or-pending uses per stage (mbarrier phase bookkeeping)
};

// Call from every thread of the CTA (blockDim.x == 256).  `raw` points at kTcSmemBytes of dynamic shared memory.
__device__ __forceinline__ void tc_ctx_init(TcGemmCtx& cx, uint8_t* raw, uint64_t* bars, uint32_t* tmem_slot) {
    cx.buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    cx.bar = bars;
    cx.uses[0] = cx.uses[1] = 0;
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc<128>(tmem_slot);
    if (threadIdx.x == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    cx.tmem = *tmem_slot;
}
__device__ __forceinline__ void tc_ctx_fini(TcGemmCtx& cx) {
    tc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<128>(cx.tmem);
}

__device__ __forceinline__ uint32_t tc_pack_bf16x2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// 8 fp32 -> one 16-byte chunk of hi and one of lo, stored at the swizzled position of (row r, chunk c)
__device__ __forceinline__ void tc_store_chunk(uint8_t* hi_base, uint8_t* lo_base, int r, int c, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e], b = v[2 * e + 1];
        const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
        h[e] = tc_pack_bf16x2(ah, bh);
        l[e] = tc_pack_bf16x2(a - ah, b - bh);
    }
    const uint32_t off = (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ (r & 7)) & 7) << 4));
    *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Stage rows [r0, r0+128) x k in [k0, k0+64) of the logical matrix X(r,k):
//   KCONTIG: X(r,k) = P[r*ld + k]   (k contiguous in memory)      else: X(r,k) = P[k*ld + r]   (r contiguous)
// rows >= R and k >= K are zero; kscale (optional) multiplies X(r,k) by kscale[k].
// Split in a load half and a store half so that a caller can put ALL global loads of a k-chunk (both operands, 64
// per thread) in flight before the first conversion / shared store: the staging is latency bound otherwise.
template <bool KCONTIG>
__device__ __forceinline__ void tc_stage_load(float (&v)[4][8], const float* __restrict__ P, int ld, int r0, int R, int k0, int K,
                                              const float* __restrict__ kscale) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int id = threadIdx.x + it * 256;            // 1024 chunks of 8 elements
        const int r = KCONTIG ? (id >> 3) : (id & 127);
        const int c = KCONTIG ? (id & 7) : (id >> 7);
        const int gr = r0 + r, gk = k0 + c * 8;
        const float* src = KCONTIG ? P + (size_t)gr * ld + gk : P + (size_t)gk * ld + gr;
        const size_t estride = KCONTIG ? 1 : (size_t)ld;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = 0.0f;
            if (gr < R && gk + e < K) {
                x = __ldg(src + e * estride);
                if (kscale) x *= __ldg(kscale + gk + e);
            }
            v[it][e] = x;
        }
    }
}
template <bool KCONTIG>
__device__ __forceinline__ void tc_stage_store(uint8_t* hi_base, uint8_t* lo_base, const float (&v)[4][8]) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int id = threadIdx.x + it * 256;
        const int r = KCONTIG ? (id >> 3) : (id & 127);
        const int c = KCONTIG ? (id & 7) : (id >> 7);
        tc_store_chunk(hi_base, lo_base, r, c, v[it]);
    }
}

// One 128 x 128 output tile at (m0, n0):  C = op(A) op(B),  !TA: A[m*lda+k], TA: A[k*lda+m];  !TB: B[k*ldb+n], TB: B[n*ldb+k].
// epi(m, n, value) is called once per valid element by the thread that owns TMEM lane m.
template <bool TA, bool TB, class Epi>
__device__ __forceinline__ void cta_gemm_tc_tile(const float* A, int lda, const float* B, int ldb, int M, int Nn, int K,
                                                 const float* kscale, int m0, int n0, TcGemmCtx& cx, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idesc = tc::umma_idesc_bf16(kTcTile, kTcTile);
    const int nchunks = (K + kTcKChunk - 1) / kTcKChunk;
    int last_stage = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        if (cx.uses[s] > 0) tc::mbar_wait(&cx.bar[s], (cx.uses[s] - 1) & 1);   // stage free again
        uint8_t* st = cx.buf + s * kTcStageBytes;
        {
            float va[4][8], vb[4][8];
            tc_stage_load<!TA>(va, A, lda, m0, M, c * kTcKChunk, K, nullptr);
            tc_stage_load<TB>(vb, B, ldb, n0, Nn, c * kTcKChunk, K, kscale);
            tc_stage_store<!TA>(st, st + kTcOperandBytes, va);
            tc_stage_store<TB>(st + 2 * kTcOperandBytes, st + 3 * kTcOperandBytes, vb);
        }
        tc::fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::tc_fence_after();
            const uint32_t a0 = tc::smem_u32(st), b0 = a0 + 2 * kTcOperandBytes;
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
    tc::mbar_wait(&cx.bar[last_stage], (cx.uses[last_stage] - 1) & 1);   // commit tracks every earlier MMA too
    tc::tc_fence_after();
    const int g = warp & 3, half = warp >> 2;
    const int m = m0 + g * 32 + lane;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        const int cb = half * 64 + ch * 32;
        tc::tmem_ld_32x32b_x32(cx.tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)cb, v);
        tc::tmem_ld_wait();
        if (m < M) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = n0 + cb + i;
                if (n < Nn) epi(m, n, v[i]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();     // TMEM may be overwritten by the next tile's first MMA; epilogue stores visible CTA-wide
}

// whole matrix, tiles walked by this one CTA (monolithic kernels)
template <bool TA, bool TB, class Epi>
__device__ __forceinline__ void cta_gemm_tc(const float* A, int lda, const float* B, int ldb, int M, int Nn, int K,
                                            const float* kscale, TcGemmCtx& cx, Epi epi) {
    for (int m0 = 0; m0 < M; m0 += kTcTile)
        for (int n0 = 0; n0 < Nn; n0 += kTcTile) cta_gemm_tc_tile<TA, TB>(A, lda, B, ldb, M, Nn, K, kscale, m0, n0, cx, epi);
}

}  // namespace crw
