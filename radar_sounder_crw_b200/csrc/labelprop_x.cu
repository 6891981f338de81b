// labelprop_x.cu -- label-propagation affinity + top-k for sm_100a: tensor-core FILTER + exact fp32 REFINE.
//
// Replaces the same reference code as labelprop_f32.cu (batched_affinity, src/imported/maskedatt.py:151-175, with the radius
// mask of :232-245) and produces BIT-IDENTICAL W / I (hence masks and labels) to the pinned-order fp32 path and to
// oracle/crw_oracle.c, at tensor-core speed (precision = CRW_PREC_TC_EXACT):
//
//   lp_prep_x_kernel    F.normalize in the pinned order -> xn (fp32) and one fp16 plane h = fp16(256 xn); per-call maxima of
//                       |xn|^2 and |xn - h/256|^2 (they size the filter margin)
//   lp_filter_kernel    ONE tcgen05 pass (kind::f16, fp16 operands, fp32 accumulate in TMEM) over the dense (query tile x key
//                       tile) blocks.  The approximate dot a~ differs from the pinned fp32 chain dot by at most E (Cauchy-
//                       Schwarz on the rounding residuals + accumulation slack), so every candidate whose a~ is within 2E of
//                       the k-th best a~ survives: the true top-k is a subset of the survivors.  Four 128-row query tiles
//                       share every 64-row key stage (the L2 -> SM key stream is a quarter of one tile per stage).
//                         warp 16      TMA producer: ring of 16 KB stages (query half-tiles, then key tiles), SWIZZLE_128B
//                         warp 17      MMA issuer: query tiles copied smem -> TMEM (tcgen05.cp), TS-form MMAs, 4 accumulators
//                         warps 0-15   epilogue: thread = query row; branch-free "beats the bound and lies in the window/band"
//                                      test per value, survivors appended to a per-thread smem column as packed 32-bit keys
//                                      (20-bit fixed-point value | 12-bit stream column); a warp-wide flush merges them into a
//                                      sorted register list of KL entries whose KT-th entry gives the bound
//   lp_refine_kernel    per query: the fp32 rows of its survivors are fetched with cp.async.bulk (512 B each, mbarrier ring),
//                       dot = the oracle's sequential fmaf chain, exact top-k (logit desc, id asc), masked fill, pinned
//                       softmax, W / I stores.  Queries whose survivor list overflowed (degenerate inputs: many exact ties)
//                       are rescanned in full by a warp -- slower, same result.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace crw {

// ------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------
constexpr float kXScale = 256.0f;            // fp16 plane holds 256 * xn: keeps small components out of the subnormal range

constexpr int kXStatSlots = 32;           // the per-call maxima are spread over 32 slots (fewer same-address atomics), reduced by the filter

__global__ void __launch_bounds__(256) lp_prep_x_kernel(const float* __restrict__ x, int64_t rows, int do_normalize,
                                                        float* __restrict__ xn, __half* __restrict__ h, unsigned* __restrict__ stats) {
    __shared__ float s_n2[8], s_e2[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    float n2 = 0.0f, e2 = 0.0f;
    if (row < rows) {
        const float* xr = x + row * 128;
        float v[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) v[m] = xr[lane + 32 * m];
        if (do_normalize) {          // pinned order: identical to l2_normalize_kernel / crw_oracle_l2_normalize
            float ss = 0.0f;
#pragma unroll
            for (int m = 0; m < 4; ++m) ss = __fmaf_rn(v[m], v[m], ss);
            ss = warp_sum_butterfly_rn(ss);
            const float d = fmaxf(__fsqrt_rn(ss), kNormEps);
#pragma unroll
            for (int m = 0; m < 4; ++m) v[m] = __fdiv_rn(v[m], d);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const __half hv = __float2half_rn(v[m] * kXScale);
            const float r = v[m] - __half2float(hv) * (1.0f / kXScale);
            n2 = fmaf(v[m], v[m], n2);
            e2 = fmaf(r, r, e2);
            h[row * 128 + lane + 32 * m] = hv;
            if (xn) xn[row * 128 + lane + 32 * m] = v[m];
        }
        // one butterfly for both sums: the pair travels together
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            n2 += __shfl_xor_sync(0xffffffffu, n2, off);
            e2 += __shfl_xor_sync(0xffffffffu, e2, off);
        }
    }
    if (lane == 0) { s_n2[warp] = n2; s_e2[warp] = e2; }
    __syncthreads();
    if (threadIdx.x == 0) {          // non-negative floats order like their bit patterns
        float a = 0.0f, b2 = 0.0f;
        for (int w = 0; w < 8; ++w) { a = fmaxf(a, s_n2[w]); b2 = fmaxf(b2, s_e2[w]); }
        atomicMax(&stats[2 * (blockIdx.x % kXStatSlots)], __float_as_uint(a));
        atomicMax(&stats[2 * (blockIdx.x % kXStatSlots) + 1], __float_as_uint(b2));
    }
}

// ------------------------------------------------------------------------------------------
// filter
// ------------------------------------------------------------------------------------------
constexpr int kXBM = 128;                    // query rows per tile (TMEM lanes)
constexpr int kXBN = 64;                     // key rows per stage (TMEM columns per accumulator)
constexpr int kXG = 4;                       // query tiles per item (share the key stream)
constexpr int kXStageBytes = 16384;          // key tile [kblock 0,1][64 rows][128 B]  or  half a query tile [128 rows][128 B]
constexpr int kXStages = 5;
constexpr int kXEpi = 16;                    // epilogue warps
constexpr int kXCap = 32;                    // appended-survivor slots per thread between two flushes
constexpr int kXThreads = (kXEpi + 2) * 32;         // 16 epilogue warps, the TMA producer, the MMA issuer
constexpr int kXColBits = 12;                // packed key: value << 12 | stream column (key tile * 64 + column)
constexpr float kXFixBias = 278528.0f;       // (dot + 1.0625) * 2^18 with a = 65536 dot:  fix = 4 a + 278528
constexpr float kXFixMagic = 12582912.0f;    // 1.5 * 2^23: float -> integer in the low mantissa bits

struct XParams {
    int R, T, N, ctx, rb, k;
    int rows_rg;                             // T * N
    int rows_per_item, items_per_rg, early_items_rg;
    int v_begin, v_end;                      // schedule slots walked by this launch
    unsigned magic_n;
    int debug;
    long long total_rows;
    int32_t* surv;                           // [KL][total_rows] radargram-relative key rows of the survivors
    int32_t* cnt;                            // [total_rows]     number of survivors, bit 30 = list overflowed (rescan in full)
    const unsigned* stats;
    float inv_temp;
};

struct XItem {
    int rg, ra, rb;                          // radargram, query rows [ra, rb)
    int n_hi, f_lo, has_f0, n_kt;
};
struct XTile {
    int active, r_lo, r_hi, nhi, flo;
};

__host__ __device__ __forceinline__ void x_slot_to_item(const XParams& p, int v, int& rg, int& it) {
    const int E = p.early_items_rg, early_total = p.R * E;
    if (v < early_total) { rg = v / E; it = v - rg * E; return; }
    const int w = v - early_total, rest = p.items_per_rg - E;
    rg = w / rest;
    it = E + (w - rg * rest);
}
__host__ __device__ __forceinline__ XItem x_item(const XParams& p, int v) {
    XItem t;
    int it;
    x_slot_to_item(p, v, t.rg, it);
    t.ra = it * p.rows_per_item;
    t.rb = min(t.ra + p.rows_per_item, p.rows_rg);
    const int n_lo = max(1, t.ra / p.N);
    t.n_hi = min(p.T - 1, (t.rb - 1) / p.N);
    t.f_lo = max(0, n_lo - p.ctx);
    t.has_f0 = (t.f_lo > 0) ? ceil_div(p.N, kXBN) : 0;
    t.n_kt = (n_lo > t.n_hi || t.ra >= t.rb) ? 0 : t.has_f0 + ceil_div((t.n_hi - t.f_lo) * p.N, kXBN);
    return t;
}
__host__ __device__ __forceinline__ XTile x_tile(const XParams& p, const XItem& t, int g) {
    XTile q;
    q.r_lo = t.ra + g * kXBM;
    q.r_hi = min(q.r_lo + kXBM, t.rb);
    q.active = q.r_lo < q.r_hi;
    const int nlo = max(1, q.r_lo / p.N);
    q.nhi = min(p.T - 1, (q.r_hi - 1) / p.N);
    if (nlo > q.nhi) q.active = 0;
    q.flo = max(0, nlo - p.ctx);
    return q;
}
// first key row (radargram-relative) and number of key rows of key tile kt of the item's stream
__host__ __device__ __forceinline__ void x_ktile_rows(const XParams& p, const XItem& t, int kt, int& row0, int& nrows) {
    if (kt < t.has_f0) { row0 = kt * kXBN; nrows = min(kXBN, p.N - row0); return; }
    row0 = t.f_lo * p.N + (kt - t.has_f0) * kXBN;
    nrows = min(kXBN, t.n_hi * p.N - row0);
}
// does query tile q look at key rows [row0, row0 + nrows)?  (its contiguous window, or frame 0)
__host__ __device__ __forceinline__ bool x_needs(const XParams& p, const XTile& q, int row0, int nrows) {
    if (!q.active) return false;
    const bool main = row0 < q.nhi * p.N && row0 + nrows > q.flo * p.N;
    const bool f0 = q.flo > 0 && row0 < p.N;
    return main || f0;
}

// kind::f16 with fp16 A / B (format 0), fp32 D
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// validity of this thread's query against the (<= 64) key rows [row0, row0 + nrows): bit c set when key row row0 + c lies in an
// allowed key frame (kf < n, and kf == 0 or kf >= win_lo) and inside the radius band.  The band of the query node is one run of
// bits (mq, first node lo_q), shifted per key frame to where that frame starts in the tile.  Needs 2 rb + 1 <= 64.
__device__ __forceinline__ unsigned long long x_band_mask(int row0, int nrows, int N, unsigned magic_n, int n, int win_lo, int lo_q,
                                                          unsigned long long mq) {
    if (nrows <= 0) return 0ull;
    const int kf0 = (int)__umulhi((unsigned)row0, magic_n);
    unsigned long long m = 0ull;
    int kf = kf0;
    for (int o = kf0 * N - row0; o < nrows; o += N, ++kf) {          // warp-uniform trip count; o = tile column of node 0
        const int sh = o + lo_q;
        const unsigned long long seg = (sh >= 0) ? ((sh < 64) ? (mq << sh) : 0ull) : ((sh > -64) ? (mq >> (-sh)) : 0ull);
        m |= ((kf < n) && (kf == 0 || kf >= win_lo)) ? seg : 0ull;
    }
    if (nrows < 64) m &= (1ull << nrows) - 1ull;
    return m;
}
// closed form for 32 <= N <= 64 (at most three frames meet a 64-row key tile): the band is the N-periodic pattern P (bits lo_q ..
// lo_q + w_q - 1) read from phase row0 mod N; the window is frame 0 plus the run of rows [win_lo N, n N)
__device__ __forceinline__ unsigned long long x_shl64(unsigned long long x, int s) { return (s >= 0 && s < 64) ? (x << s) : 0ull; }
__device__ __forceinline__ unsigned long long x_run64(int lo, int hi) {      // bits [lo, hi) of a 64-bit word, any lo / hi
    lo = max(lo, 0);
    hi = min(hi, 64);
    if (lo >= hi) return 0ull;
    const unsigned long long upto = (hi >= 64) ? ~0ull : ((1ull << hi) - 1ull);
    return upto & ~((1ull << lo) - 1ull);
}
__device__ __forceinline__ unsigned long long x_band_mask_fast(int row0, int nrows, int N, unsigned magic_n, int n, int win_lo, int lo_q,
                                                               unsigned long long mq) {
    const int kf0 = (int)__umulhi((unsigned)row0, magic_n), phi = row0 - kf0 * N;
    const unsigned long long P = mq << lo_q;                                  // lo_q + w_q <= N <= 64
    const unsigned long long band = (P >> phi) | x_shl64(P, N - phi) | x_shl64(P, 2 * N - phi);
    const unsigned long long win = x_run64(-row0, N - row0) | x_run64(win_lo * N - row0, n * N - row0);
    return band & win & x_run64(0, nrows);
}
// generic form (any band width)
__device__ __forceinline__ unsigned long long x_band_mask_wide(int row0, int nrows, int N, unsigned magic_n, int n, int win_lo, int q, int rb) {
    unsigned long long m = 0ull;
    const int kf0 = (int)__umulhi((unsigned)row0, magic_n);
    int c = 0, kf = kf0, j = row0 - kf0 * N;
    while (c < nrows) {
        const int seg = min(nrows - c, N - j);
        if ((kf < n) && (kf == 0 || kf >= win_lo)) {
            const int lo = max(j, q - rb), hi = min(j + seg - 1, q + rb);
            if (lo <= hi) {
                const int b0 = c + lo - j, nb = hi - lo + 1;
                m |= ((nb >= 64) ? ~0ull : ((1ull << nb) - 1ull)) << b0;
            }
        }
        c += seg; j = 0; ++kf;
    }
    return m;
}

// profiling aid (CRW_TC_DEBUG bit 3 = 8): cycles per phase, summed per (CTA, warp).  [kernel 0 = filter, 1 = refine][CTA][warp 0..17][8]
//   filter epilogue: [0] wait for an accumulator [1] validity mask [2] TMEM loads + scan [3] flushes [4] final list -> global [5] item total
//   refine:          [0] metadata staging [1] wait (producer: empty slot; consumer: full slot) [2] producer issue / consumer chain
//                    [3] rank + finish [4] full rescans [5] chunk total
__device__ unsigned long long g_x_prof[2 * 160 * 18 * 8];
#define XPROF(kern, idx, cyc) do { if (prof && lane == 0) atomicAdd(&g_x_prof[(((kern) * 160 + (blockIdx.x % 160)) * 18 + warp) * 8 + (idx)], (unsigned long long)(cyc)); } while (0)

template <int KL>
__device__ __forceinline__ void x_list_insert(uint32_t (&L)[KL], uint32_t P) {
#pragma unroll
    for (int s = KL - 1; s >= 1; --s) L[s] = min(L[s - 1], max(L[s], P));
    L[0] = max(L[0], P);
}

// bound of the hot loop, in accumulator units (65536 * dot), from the packed KT-th best entry and the margin
__device__ __forceinline__ float x_threshold(uint32_t theta, int m_fix) {
    if (theta == 0u) return -INFINITY;
    const int X = (int)(theta >> kXColBits) - m_fix - 1;             // one fixed-point step of slack for the rounding of fix()
    return ((float)max(X, 0) - kXFixBias) * 0.25f;
}

// 16 columns of a key tile: branch-free append of the values that beat the bound and are valid (window x band).
// Per value: validity bit -> predicate, compare, column id, predicated 8-byte store (raw value, column), predicated slot advance.
// Written as PTX so that the per-value work stays at these five instructions (the fixed-point packing happens at flush time).
template <int I, int END>
struct XScan {
    static __device__ __forceinline__ void run(const float (&val)[32], uint32_t vbits, uint32_t colbase, float thr, uint32_t& ptr) {
        asm volatile(
            "{\n\t.reg .pred p, v;\n\t.reg .b32 t, c;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 v, t, 0;\n\t"
            "setp.ge.and.f32 p, %3, %4, v;\n\t"
            "or.b32 c, %5, %6;\n\t"
            "@p st.shared.v2.b32 [%0], {%7, c};\n\t"
            "@p add.u32 %0, %0, 256;\n\t}"
            : "+r"(ptr)
            : "r"(vbits), "n"(1u << I), "f"(val[I]), "f"(thr), "r"(colbase), "n"(I), "r"(__float_as_uint(val[I]))
            : "memory");
        XScan<I + 1, END>::run(val, vbits, colbase, thr, ptr);
    }
};
template <int END>
struct XScan<END, END> {
    static __device__ __forceinline__ void run(const float (&)[32], uint32_t, uint32_t, float, uint32_t&) {}
};
template <int I0>
__device__ __forceinline__ void x_scan16(const float (&val)[32], uint32_t vbits, uint32_t colbase, float thr, uint32_t& ptr) {
    XScan<I0, I0 + 16>::run(val, vbits, colbase, thr, ptr);
}

template <int KT, int KL>
__global__ void __launch_bounds__(kXThreads, 1)
lp_filter_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap, XParams p) {
    constexpr int kProducerWarp = kXEpi, kMmaWarp = kXEpi + 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                                              // kXStages x 16 KB ring
    uint2* sApp = reinterpret_cast<uint2*>(smem + kXStages * kXStageBytes);         // [16 warps][kXCap][32 lanes] (value, column)
    __shared__ uint64_t k_full[kXStages], k_empty[kXStages], acc_full[kXG], acc_empty[kXG];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;

    if (warp == kMmaWarp) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        for (int s = 0; s < kXStages; ++s) { tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); }
        for (int g = 0; g < kXG; ++g) { tc::mbar_init(&acc_full[g], 1); tc::mbar_init(&acc_empty[g], 4); }
        tc::fence_barrier_init();
    }
    if (warp == kProducerWarp && lane == 0) { tc::prefetch_tmap(&qmap); tc::prefetch_tmap(&kmap); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;      // columns [0,256): four query tiles (64 each); [256,512): four accumulators

    if (warp == kProducerWarp) {
        // ================= TMA producer =================
        const bool leader = tc::elect_one();
        uint32_t scnt = 0;
        for (int v = p.v_begin + blockIdx.x; v < p.v_end; v += gridDim.x) {
            const XItem t = x_item(p, v);
            if (t.n_kt == 0) continue;
            const int grow = t.rg * p.rows_rg;
            for (int g = 0; g < kXG; ++g) {
                const XTile q = x_tile(p, t, g);
                if (!q.active) continue;
                for (int kb = 0; kb < 2; ++kb, ++scnt) {
                    const int s = scnt % kXStages;
                    tc::mbar_wait_backoff(&k_empty[s], ((scnt / kXStages) & 1) ^ 1);
                    if (leader) {
                        tc::mbar_arrive_expect_tx(&k_full[s], kXStageBytes);
                        tc::tma_load_2d(sK + s * kXStageBytes, &qmap, kb * 64, grow + q.r_lo, &k_full[s]);
                    }
                }
            }
            for (int kt = 0; kt < t.n_kt; ++kt, ++scnt) {
                const int s = scnt % kXStages;
                int row0, nrows;
                x_ktile_rows(p, t, kt, row0, nrows);
                tc::mbar_wait_backoff(&k_empty[s], ((scnt / kXStages) & 1) ^ 1);
                if (leader) {
                    tc::mbar_arrive_expect_tx(&k_full[s], kXStageBytes);
                    uint8_t* dst = sK + s * kXStageBytes;
                    tc::tma_load_2d(dst, &kmap, 0, grow + row0, &k_full[s]);
                    tc::tma_load_2d(dst + kXBN * 128, &kmap, 64, grow + row0, &k_full[s]);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        const bool leader = tc::elect_one();
        const uint64_t kdesc0 = tc::umma_smem_desc_k128(tc::smem_u32(sK));
        uint32_t scnt = 0, ucnt[kXG] = {0u, 0u, 0u, 0u};
        for (int v = p.v_begin + blockIdx.x; v < p.v_end; v += gridDim.x) {
            const XItem t = x_item(p, v);
            if (t.n_kt == 0) continue;
            XTile qt[kXG];
#pragma unroll
            for (int g = 0; g < kXG; ++g) qt[g] = x_tile(p, t, g);
            // query tiles: ring slots -> TMEM (tcgen05.cp executes in order with the MMAs issued before and after it)
#pragma unroll
            for (int g = 0; g < kXG; ++g) {
                if (!qt[g].active) continue;
                for (int kb = 0; kb < 2; ++kb, ++scnt) {
                    const int s = scnt % kXStages;
                    tc::mbar_wait(&k_full[s], (scnt / kXStages) & 1);
                    tc::tc_fence_after();
                    if (leader) {
                        const uint64_t sd = kdesc0 + (uint64_t)((s * kXStageBytes) >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc::tmem_cp_128x256b(tmem_base + (uint32_t)(g * 64 + kb * 32 + ks * 8), sd + (uint64_t)((ks * 32) >> 4));
                        tc::umma_commit(&k_empty[s]);
                    }
                    __syncwarp();
                }
            }
            for (int kt = 0; kt < t.n_kt; ++kt, ++scnt) {
                const int s = scnt % kXStages;
                int row0, nrows;
                x_ktile_rows(p, t, kt, row0, nrows);
                const int ncols = min(kXBN, (nrows + 15) & ~15);
                tc::mbar_wait(&k_full[s], (scnt / kXStages) & 1);
                const uint32_t idesc = umma_idesc_f16(kXBM, ncols);
                const uint64_t kdesc = kdesc0 + (uint64_t)((s * kXStageBytes) >> 4);
#pragma unroll
                for (int g = 0; g < kXG; ++g) {
                    if (!x_needs(p, qt[g], row0, nrows)) continue;
                    tc::mbar_wait(&acc_empty[g], ((ucnt[g]) & 1) ^ 1);
                    tc::tc_fence_after();
                    if (leader) {
                        const uint32_t d = tmem_base + 256u + (uint32_t)(g * kXBN);
                        if (!(p.debug & 4))
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc::umma_bf16_ts(d, tmem_base + (uint32_t)(g * 64 + kb * 32 + ks * 8),
                                                 kdesc + (uint64_t)(((kb * (kXBN * 128)) + ks * 32) >> 4), idesc, (kb | ks) ? 1u : 0u);
                        tc::umma_commit(&acc_full[g]);
                    }
                    __syncwarp();
                    ++ucnt[g];
                }
                if (leader) tc::umma_commit(&k_empty[s]);
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: thread = query row =================
        const int g = warp >> 2, quarter = warp & 3;
        const int rb = p.rb, ctx = p.ctx;
        const uint32_t app = tc::smem_u32(sApp + (size_t)warp * kXCap * 32) + lane * 8;
        // margin from the call's maxima (see header): E = 2 eps |x| + accumulation slack, survivors within 2 E of the bound
        float n2 = 0.0f, e2 = 0.0f;
        for (int i = 0; i < kXStatSlots; ++i) { n2 = fmaxf(n2, __uint_as_float(p.stats[2 * i])); e2 = fmaxf(e2, __uint_as_float(p.stats[2 * i + 1])); }
        const float nrm = sqrtf(n2) * 1.00001f, eps = sqrtf(e2) * 1.0001f + 1e-9f;
        const float E = (2.0f * eps * nrm + eps * eps + 2.5e-5f * nrm * nrm) * 1.02f;
        const int m_fix = (int)ceilf(2.0f * E * 262144.0f) + 2;
        const bool range_bad = !(n2 <= 1.002f);                    // un-normalised input: the fixed-point range does not hold
        uint32_t ucnt = 0;
        for (int v = p.v_begin + blockIdx.x; v < p.v_end; v += gridDim.x) {
            const XItem t = x_item(p, v);
            if (t.n_kt == 0) {
                // rows of frame 0 only (or nothing): no survivors
                for (int r = t.ra + tid; r < t.rb; r += kXEpi * 32) p.cnt[(size_t)t.rg * p.rows_rg + r] = 0;
                continue;
            }
            const XTile qt = x_tile(p, t, g);
            const int row = qt.r_lo + quarter * 32 + lane;
            const int n = row / N, q = row - n * N;
            const bool in_item = row < t.rb;
            const bool qvalid = qt.active && in_item && (n >= 1) && (n < p.T);
            const int win_lo = (n > ctx + 1) ? n - ctx : 1;
            const int lo_q = max(0, q - rb), w_q = min(N - 1, q + rb) - lo_q + 1;
            const unsigned long long mq = (w_q >= 64) ? ~0ull : ((1ull << w_q) - 1ull);
            const bool wide_band = 2 * rb + 1 > 64;
            const bool fast_mask = N >= 32 && N <= 64;
            uint32_t L[KL];
#pragma unroll
            for (int s = 0; s < KL; ++s) L[s] = 0u;
            uint32_t ptr = app;                                   // next free append slot of this thread (slots are 256 B apart)
            float thr = -INFINITY;
            auto flush = [&]() {
                const int cnt = (int)((ptr - app) >> 8);
                const int maxc = __reduce_max_sync(0xffffffffu, cnt);
                for (int i = 0; i < maxc; ++i) {
                    uint32_t vb, col;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(vb), "=r"(col) : "r"(app + i * 256) : "memory");
                    const float tf = __fmaf_rn(__uint_as_float(vb), 4.0f, kXFixBias + kXFixMagic);
                    const uint32_t P = (__float_as_uint(tf) << kXColBits) | col;
                    x_list_insert<KL>(L, (i < cnt) ? P : 0u);
                }
                ptr = app;
                thr = x_threshold(L[KT - 1], m_fix);
            };
            auto maybe_flush = [&]() { if (__any_sync(0xffffffffu, ptr > app + (kXCap - 16) * 256)) flush(); };
            for (int kt = 0; kt < t.n_kt; ++kt) {
                int row0, nrows;
                x_ktile_rows(p, t, kt, row0, nrows);
                if (!x_needs(p, qt, row0, nrows)) continue;                      // warp-uniform (same test as the MMA warp)
                unsigned long long vm = 0ull;
                if (qvalid) vm = wide_band ? x_band_mask_wide(row0, nrows, N, p.magic_n, n, win_lo, q, rb)
                                 : (fast_mask ? x_band_mask_fast(row0, nrows, N, p.magic_n, n, win_lo, lo_q, mq)
                                              : x_band_mask(row0, nrows, N, p.magic_n, n, win_lo, lo_q, mq));
                if (p.debug & 1) vm = 0ull;
                tc::mbar_wait(&acc_full[g], ucnt & 1);
                tc::tc_fence_after();
                ++ucnt;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + (uint32_t)(g * kXBN);
                float val[32];
                if (!(p.debug & 2)) {
                    tc::tmem_ld_32x32b_x32(taddr, val);
                    tc::tmem_ld_wait();
                    x_scan16<0>(val, (uint32_t)vm, (uint32_t)(kt * 64), thr, ptr);
                    maybe_flush();
                    x_scan16<16>(val, (uint32_t)vm, (uint32_t)(kt * 64), thr, ptr);
                    maybe_flush();
                }
                if (nrows > 32 && !(p.debug & 2)) {
                    tc::tmem_ld_32x32b_x32(taddr + 32u, val);
                    tc::tmem_ld_wait();
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[g]);
                if (nrows > 32 && !(p.debug & 2)) {
                    x_scan16<0>(val, (uint32_t)(vm >> 32), (uint32_t)(kt * 64 + 32), thr, ptr);
                    maybe_flush();
                    x_scan16<16>(val, (uint32_t)(vm >> 32), (uint32_t)(kt * 64 + 32), thr, ptr);
                    maybe_flush();
                }
            }
            flush();
            // survivors: entries within the margin of the KT-th best (all of them when fewer than KT exist)
            if (qt.r_lo < qt.r_hi && in_item) {
                const size_t grow = (size_t)t.rg * p.rows_rg + row;
                int c = 0;
                bool ovf = range_bad;
                if (qvalid) {
                    const uint32_t theta = L[KT - 1];
                    const uint32_t mP = (uint32_t)m_fix << kXColBits;
                    const uint32_t bound = (theta > mP) ? theta - mP : 0u;
                    if (theta != 0u && L[KL - 1] != 0u && L[KL - 1] >= bound) ovf = true;      // more may lie within the margin
#pragma unroll
                    for (int s = 0; s < KL; ++s) {
                        if (L[s] != 0u && L[s] >= bound) {
                            const int col = (int)(L[s] & ((1u << kXColBits) - 1u));
                            const int kt = col >> 6, i = col & 63;
                            const int kr = (kt < t.has_f0) ? kt * kXBN + i : t.f_lo * N + (kt - t.has_f0) * kXBN + i;
                            p.surv[(size_t)c * p.total_rows + grow] = kr;
                            ++c;
                        }
                    }
                }
                p.cnt[grow] = qvalid ? (c | (ovf ? (1 << 30) : 0)) : 0;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// refine
// ------------------------------------------------------------------------------------------
constexpr int kRRowBytes = 528;              // 128 floats + 16 B pad: float4 row reads of 32 lanes are conflict-free
constexpr int kRChunk = 128;                 // queries per metadata chunk
constexpr int kRMaxWarps = 16;

struct RParams {
    const float* xn;         // [total_rows, 128] normalised features
    const int32_t* surv;     // [KL][total_rows]
    const int32_t* cnt;      // [total_rows]
    float* W;                // [R, T, k, N]
    int32_t* I;
    int R, T, N, ctx, rb, k;
    int rows_rg, row_begin, row_end;   // radargram-relative row range handled by this launch (every radargram)
    long long total_rows;
    float inv_temp;
    unsigned magic_n;
    int debug;               // timing aids (CRW_TC_DEBUG): 16 = no W / I stores, 32 = no row copies, 64 = no chains (results invalid)
};

// one 512-byte feature row global -> shared by the whole warp (16 bytes per lane, asynchronous: cp.async / LDGSTS)
__device__ __forceinline__ void warp_row_copy_async(uint32_t dst_row_smem, const float* src_row, int lane) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_row_smem + lane * 16), "l"(src_row + lane * 4) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// candidate id (slot in the trimmed key set * N + node) of key row kr for query frame n
__device__ __forceinline__ int x_cand_id(int kr, int n, int N, int ctx, unsigned magic_n) {
    const int kf = (int)__umulhi((unsigned)kr, magic_n), j = kr - kf * N;
    const int slot = (n > ctx + 1 && kf != 0) ? kf - (n - ctx) + 1 : kf;
    return slot * N + j;
}

// the oracle's dot product: sequential fmaf chain over the 128 channels of two staged rows
__device__ __forceinline__ float x_chain_dot(const uint8_t* krow_smem, const uint8_t* qrow_smem) {
    const float4* krow = reinterpret_cast<const float4*>(krow_smem);
    const float4* qrow = reinterpret_cast<const float4*>(qrow_smem);
    float acc = 0.0f;
#pragma unroll 8
    for (int c4 = 0; c4 < 32; ++c4) {
        const float4 kv = krow[c4], qv = qrow[c4];
        acc = __fmaf_rn(kv.x, qv.x, acc);
        acc = __fmaf_rn(kv.y, qv.y, acc);
        acc = __fmaf_rn(kv.z, qv.z, acc);
        acc = __fmaf_rn(kv.w, qv.w, acc);
    }
    return acc;
}

// Full exact scan of one query by one warp (its survivor list overflowed: many candidates tie within the filter margin).  The
// in-band candidates are visited in ascending id order, 32 at a time through the warp's staging buffer; lane i keeps the i-th
// best (the insertion of lp_topk_f32_kernel).  Returns this lane's (logit, id).
// round_begin / round_stride: this warp visits rounds round_begin, round_begin + round_stride, ... (0, 1 = all of them)
template <int CAP>
__device__ __forceinline__ void x_full_scan(const RParams& p, const float* xr, uint8_t* buf, int n, int q, int k, int round_begin, int round_stride,
                                            float& v_out, int& id_out) {
    constexpr int kStep = CAP < 32 ? CAP : 32;      // candidates per round
    const int lane = threadIdx.x & 31;
    const int N = p.N, rb = p.rb;
    const int F = n_key_frames(n, p.ctx);
    const unsigned kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    float v = -INFINITY;
    int id = 0;
    float thr = -INFINITY;
    const int jlo = max(0, q - rb), jhi = min(N - 1, q + rb), bw = jhi - jlo + 1;
    const uint32_t sbuf = tc::smem_u32(buf);
    warp_row_copy_async(sbuf + CAP * kRRowBytes, xr + ((size_t)n * N + q) * 128, lane);
    const int total = F * bw;                          // in-band candidates, ascending id = (frame slot, node)
    for (int c0 = round_begin * kStep; c0 < total; c0 += round_stride * kStep) {
        const int nrow = min(kStep, total - c0);
        __syncwarp();
        for (int r = 0; r < nrow; ++r) {
            const int cc = c0 + r, f = cc / bw, jj = cc - f * bw;
            warp_row_copy_async(sbuf + r * kRRowBytes, xr + ((size_t)key_frame(n, p.ctx, f) * N + jlo + jj) * 128, lane);
        }
        cp_async_wait_all();
        __syncwarp();
        const bool ok = lane < nrow;
        float cand = -INFINITY;
        if (ok) cand = __fmul_rn(x_chain_dot(buf + lane * kRRowBytes, buf + CAP * kRRowBytes), p.inv_temp);
        const int cc = c0 + lane, fl = cc / bw;
        const int cid = fl * N + jlo + (cc - fl * bw);
        unsigned m = __ballot_sync(0xffffffffu, ok && cand > thr);
        while (m) {
            const int s = __ffs(m) - 1;
            m &= m - 1;
            const float c = __shfl_sync(0xffffffffu, cand, s);
            const int ci = __shfl_sync(0xffffffffu, cid, s);
            if (c > thr) {
                const int pos = __popc(__ballot_sync(0xffffffffu, v >= c) & kmask);
                const float vup = __shfl_up_sync(0xffffffffu, v, 1);
                const int iup = __shfl_up_sync(0xffffffffu, id, 1);
                if (lane > pos) { v = vup; id = iup; }
                else if (lane == pos) { v = c; id = ci; }
                if (lane >= k) v = -INFINITY;
                thr = __shfl_sync(0xffffffffu, v, k - 1);
            }
        }
    }
    __syncwarp();
    v_out = v;
    id_out = id;
}

// Per lane group (sub_base .. sub_base + SL): lane sl holds the winner of rank sl, (logit desc, id asc), `live` of them real.
// Masked fill, pinned softmax (sequential sum in rank order), W / I stores.  Every lane of the warp calls this; groups without a
// query pass active = false.  es: k floats of scratch per group.
__device__ __forceinline__ void x_finish_group(const RParams& p, float v, int id, bool active, int live, int rg, int n, int q, int k,
                                               int sub_base, int sl, float* es) {
    const int N = p.N, rb = p.rb;
    if (active && sl >= live && sl < k) {
        // fewer than k in-band candidates: out-of-band ones share one logit and come in ascending id order
        const int lo_q = max(0, q - rb), w_q = min(N - 1, q + rb) - lo_q + 1, nob = max(N - w_q, 1);
        const int t = sl - live, f = t / nob, r = t - f * nob;
        const int jj = (r < lo_q) ? r : r + w_q;
        v = __fmul_rn(kMaskBias, p.inv_temp);
        id = f * N + jj;
    }
    const float v0 = __shfl_sync(0xffffffffu, v, sub_base);
    const bool w = active && sl < k;
    const float e = w ? pinned_expf(__fsub_rn(v, v0)) : 0.0f;
    if (w) es[sl] = e;
    __syncwarp();
    float ssum = 1.0f;
    if (w) {
        ssum = es[0];
        for (int j = 1; j < k; ++j) ssum = __fadd_rn(ssum, es[j]);
    }
    __syncwarp();
    if (w && !(p.debug & 16)) {
        const size_t o = ((size_t)(rg * p.T + n) * k + sl) * N + q;
        p.W[o] = __fdiv_rn(e, ssum);
        p.I[o] = id;
    }
}

// SL = survivor slots per query handled by one lane group (16: two queries per pass; 32: one query per pass); WARPS autonomous
// warps per CTA, each with a staging buffer of SROWS feature rows (the last 32 / SL of them hold the query rows).
// A CTA owns a contiguous range of query rows and walks it in chunks of kRChunk queries: all threads first stage the chunk's
// survivor lists (counts + key rows) in shared memory with coalesced loads.  Then every warp works on its own passes: the fp32
// rows of the survivors (and the query rows) are copied into the warp's staging buffer with cp.async, one coalesced 512-byte
// row per instruction, packed in survivor order; each lane then runs the sequential chain for its own survivor slot from
// shared memory.  The bytes in flight per SM (what bounds this gather) are WARPS x ~11 rows x 512 B.
template <int SL, int WARPS, int SROWS>
__global__ void __launch_bounds__(WARPS * 32, 1) lp_refine_kernel(RParams p) {
    constexpr int QPB = 32 / SL;             // queries per pass
    constexpr int kCapRows = SROWS - QPB;    // survivor rows one pass can stage
    constexpr int kBufBytes = SROWS * kRRowBytes;
    constexpr int kThreads = WARPS * 32;
    extern __shared__ __align__(128) uint8_t rsm[];
    __shared__ float es_all[WARPS][2][32];
    __shared__ float sc_v[WARPS][32];
    __shared__ int sc_i[WARPS][32];
    __shared__ int s_cnt[kRChunk];
    constexpr int kMaxOvf = 64;              // overflowed queries of a chunk that are rescanned by the whole CTA (more: inline, one warp each)
    __shared__ int s_ovf[kMaxOvf], s_novf;
    int* s_kr = reinterpret_cast<int*>(rsm + (size_t)WARPS * kBufBytes);      // [SL][kRChunk]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, k = p.k;
    // this CTA's share of the launch: rows [r_lo, r_hi) of the flattened (radargram, row in [row_begin, row_end)) space
    const int rows_launch = p.row_end - p.row_begin;
    const long long total_q = (long long)p.R * rows_launch;
    long long per = (total_q + gridDim.x - 1) / gridDim.x;
    per = (per + QPB - 1) / QPB * QPB;
    const long long r_lo = (long long)blockIdx.x * per, r_hi = min(total_q, r_lo + per);
    const int ql = lane / SL, s = lane % SL, sub_base = ql * SL;
    const bool prof = (p.debug & 8) != 0;
    uint8_t* buf = rsm + (size_t)warp * kBufBytes;
    const uint32_t sbuf = tc::smem_u32(buf);

    for (long long c0 = r_lo; c0 < r_hi; c0 += kRChunk) {
        const int nq = (int)min((long long)kRChunk, r_hi - c0);
        const long long c_s0 = prof ? clock64() : 0;
        if (tid == 0) s_novf = 0;
        // ---- stage the chunk's metadata: every thread issues its loads back to back (independent, coalesced along the rows) ----
        const int rg0 = (int)(c0 / rows_launch);
        const int row0 = p.row_begin + (int)(c0 - (long long)rg0 * rows_launch);        // chunk row i: row0 + i, wrapping into the next radargram
        {
            constexpr int kPer = (SL * kRChunk + kThreads - 1) / kThreads;
            int vals[kPer];
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int j = tid + u * kThreads, sl = j / kRChunk, i = j - sl * kRChunk;
                int rg = rg0, row = row0 + i;
                while (row >= p.row_end) { row -= rows_launch; ++rg; }
                vals[u] = (sl < SL && i < nq) ? __ldg(p.surv + (size_t)sl * p.total_rows + (size_t)rg * p.rows_rg + row) : 0;
            }
            if (tid < nq) {
                int rg = rg0, row = row0 + tid;
                while (row >= p.row_end) { row -= rows_launch; ++rg; }
                s_cnt[tid] = __ldg(p.cnt + (size_t)rg * p.rows_rg + row);
            }
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int j = tid + u * kThreads;
                if (j < SL * kRChunk) s_kr[j] = vals[u];
            }
        }
        __syncthreads();
        const long long c_s1 = prof ? clock64() : 0;
        XPROF(1, 0, c_s1 - c_s0);
        const int nbuf = (nq + QPB - 1) / QPB;
        for (int b = warp; b < nbuf; b += WARPS) {
            const int i = b * QPB + ql;
            int rg = rg0, row = row0 + i;
            while (row >= p.row_end) { row -= rows_launch; ++rg; }
            const int n = row / N, q = row - n * N;
            int c = (i < nq) ? s_cnt[i] : 0;
            const bool rescan = (c >> 30) & 1;
            c = rescan ? 0 : (c & 0xffff);
            const bool valid = s < c;
            const int kr = valid ? s_kr[s * kRChunk + i] : 0;
            const bool is_query = i < nq && n >= 1 && n < p.T;
            const unsigned vall = __ballot_sync(0xffffffffu, valid);
            // the staging buffer holds kCapRows survivor rows: when the pass has more (two long lists), the lane groups go one by one
            const int nsteps = (__popc(vall) > kCapRows) ? QPB : 1;
            for (int step = 0; step < nsteps; ++step) {
                const long long c_c0 = prof ? clock64() : 0;
                const bool mine = valid && (nsteps == 1 || ql == step);
                const unsigned vmask = __ballot_sync(0xffffffffu, mine);
                const int srow = __popc(vmask & ((1u << lane) - 1u));          // this lane's row of the staging buffer
                if (!(p.debug & 32)) {
                    unsigned m = vmask;
                    int r_stage = 0;
                    while (m) {
                        const int r = __ffs(m) - 1;
                        m &= m - 1;
                        const int kr_r = __shfl_sync(0xffffffffu, kr, r);
                        const int rg_r = __shfl_sync(0xffffffffu, rg, r);
                        warp_row_copy_async(sbuf + r_stage * kRRowBytes, p.xn + ((size_t)rg_r * p.rows_rg + kr_r) * 128, lane);
                        ++r_stage;
                    }
#pragma unroll
                    for (int grp = 0; grp < QPB; ++grp) {
                        if (nsteps > 1 && grp != step) continue;
                        const int c_g = __shfl_sync(0xffffffffu, c, grp * SL);
                        if (c_g > 0) {
                            const int rg_g = __shfl_sync(0xffffffffu, rg, grp * SL), row_g = __shfl_sync(0xffffffffu, row, grp * SL);
                            warp_row_copy_async(sbuf + (kCapRows + grp) * kRRowBytes, p.xn + ((size_t)rg_g * p.rows_rg + row_g) * 128, lane);
                        }
                    }
                }
                cp_async_wait_all();
                __syncwarp();
                const long long c_c1 = prof ? clock64() : 0;
                float acc = 0.0f;
                if (mine && !(p.debug & 64)) acc = x_chain_dot(buf + srow * kRRowBytes, buf + (kCapRows + ql) * kRRowBytes);
                __syncwarp();                                                 // the staging buffer may be refilled
                const long long c_c2 = prof ? clock64() : 0;
                const float lg = mine ? __fmul_rn(acc, p.inv_temp) : -INFINITY;
                const int id = mine ? x_cand_id(kr, n, N, p.ctx, p.magic_n) : 0x7fffffff;
                // rank among the query's survivors: (logit desc, id asc)
                int rank = 0;
#pragma unroll
                for (int o = 0; o < SL; ++o) {
                    const float lo_ = __shfl_sync(0xffffffffu, lg, sub_base + o);
                    const int io = __shfl_sync(0xffffffffu, id, sub_base + o);
                    rank += (lo_ > lg || (lo_ == lg && io < id)) ? 1 : 0;
                }
                // move every survivor to the lane of its rank within the query's lane group (rank < c <= SL)
                if (mine) { sc_v[warp][sub_base + rank] = lg; sc_i[warp][sub_base + rank] = id; }
                __syncwarp();
                float v = -INFINITY;
                int idv = 0;
                const bool grp_on = nsteps == 1 || ql == step;
                if (grp_on && s < c) { v = sc_v[warp][lane]; idv = sc_i[warp][lane]; }
                __syncwarp();
                x_finish_group(p, v, idv, grp_on && is_query && !rescan, min(c, k), rg, n, q, k, sub_base, s, es_all[warp][ql]);
                if (prof) { const long long c_c3 = clock64(); XPROF(1, 1, c_c1 - c_c0); XPROF(1, 2, c_c2 - c_c1); XPROF(1, 3, c_c3 - c_c2); }
            }
            // overflowed lists: queued for a rescan by the whole CTA after the chunk (inline by this warp when the queue is full)
            const long long c_r0 = prof ? clock64() : 0;
#pragma unroll
            for (int grp = 0; grp < QPB; ++grp) {
                const int need = __shfl_sync(0xffffffffu, (rescan && is_query) ? 1 : 0, grp * SL);
                if (!need) continue;                                         // warp-uniform
                const int g_i = __shfl_sync(0xffffffffu, i, grp * SL);
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&s_novf, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot < kMaxOvf) {
                    if (lane == 0) s_ovf[slot] = g_i;
                    continue;
                }
                const int g_n = __shfl_sync(0xffffffffu, n, grp * SL), g_q = __shfl_sync(0xffffffffu, q, grp * SL);
                const int g_rg = __shfl_sync(0xffffffffu, rg, grp * SL);
                float fv;
                int fid;
                x_full_scan<kCapRows>(p, p.xn + (size_t)g_rg * p.rows_rg * 128, buf, g_n, g_q, k, 0, 1, fv, fid);
                const int g_live = __popc(__ballot_sync(0xffffffffu, fv > -INFINITY) & ((k >= 32) ? 0xffffffffu : ((1u << k) - 1u)));
                x_finish_group(p, fv, fid, true, g_live, g_rg, g_n, g_q, k, 0, lane, es_all[warp][0]);
            }
            if (prof) XPROF(1, 4, clock64() - c_r0);
        }
        __syncthreads();
        // ---- queued rescans: every warp scans its share of the candidate rounds, warp 0 merges the partial lists ----
        {
            const long long c_r0 = prof ? clock64() : 0;
            const int novf = min(s_novf, kMaxOvf);
            const unsigned kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
            for (int e = 0; e < novf; ++e) {
                const int i = s_ovf[e];
                int rg = rg0, row = row0 + i;
                while (row >= p.row_end) { row -= rows_launch; ++rg; }
                const int n = row / N, q = row - n * N;
                float pv;
                int pi;
                x_full_scan<kCapRows>(p, p.xn + (size_t)rg * p.rows_rg * 128, buf, n, q, k, warp, WARPS, pv, pi);
                sc_v[warp][lane] = pv;
                sc_i[warp][lane] = pi;
                __syncthreads();
                if (warp == 0) {
                    float v = -INFINITY;
                    int id = 0;
                    for (int w2 = 0; w2 < WARPS; ++w2)
                        for (int s2 = 0; s2 < k; ++s2) {
                            const float c = sc_v[w2][s2];
                            const int ci = sc_i[w2][s2];
                            if (!(c > -INFINITY)) break;                                  // (warp-uniform) lists are sorted: the rest is empty
                            // candidates arrive in no particular order: full comparator (logit desc, id asc)
                            const int pos = __popc(__ballot_sync(0xffffffffu, v > c || (v == c && id < ci)) & kmask);
                            if (pos < k) {
                                const float vup = __shfl_up_sync(0xffffffffu, v, 1);
                                const int iup = __shfl_up_sync(0xffffffffu, id, 1);
                                if (lane > pos) { v = vup; id = iup; }
                                else if (lane == pos) { v = c; id = ci; }
                                if (lane >= k) v = -INFINITY;
                            }
                        }
                    const int live = __popc(__ballot_sync(0xffffffffu, v > -INFINITY) & kmask);
                    x_finish_group(p, v, id, true, live, rg, n, q, k, 0, lane, es_all[0][0]);
                }
                __syncthreads();
            }
            if (prof) XPROF(1, 4, clock64() - c_r0);
        }
        __syncthreads();                      // the metadata of this chunk is dead
        if (prof) XPROF(1, 5, clock64() - c_s0);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct LpXPlan {
    alignas(64) unsigned char maps[2 * sizeof(CUtensorMap)];
    XParams p;
    RParams r;
    int kt, kl;
};

static void x_kt_kl(int k, int& kt, int& kl) {
    if (k <= 10) { kt = 10; kl = 16; }
    else if (k <= 16) { kt = 16; kl = 24; }
    else { kt = 24; kl = 32; }
}
int lp_x_max_k() { return 24; }

// scratch layout of the exact tensor path: stats | fp16 plane | xn (when normalising) | survivors | counts
size_t lp_x_scratch_bytes(int R, int T, int N, int C, int k, int do_normalize) {
    int kt, kl;
    x_kt_kl(k, kt, kl);
    const size_t rows = (size_t)R * T * N;
    size_t b = 256;
    b += align_up(rows * C * sizeof(__half), 256);
    if (do_normalize) b += align_up(rows * C * sizeof(float), 256);
    b += align_up(rows * kl * sizeof(int32_t), 256);
    b += align_up(rows * sizeof(int32_t), 256);
    return b;
}

template <int KT, int KL>
static int launch_filter(const LpXPlan& plan, const XParams& p, int max_ctas, cudaStream_t st) {
    static size_t smem_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t smem = (dev >= 0 && dev < 64) ? smem_dev[dev] : 0;
    if (smem == 0) {
        cudaFuncAttributes fa;
        CRW_CUDA_RET(cudaFuncGetAttributes(&fa, lp_filter_kernel<KT, KL>));
        const size_t slack = (1024 - (fa.sharedSizeBytes % 1024)) % 1024;
        smem = slack + (size_t)kXStages * kXStageBytes + (size_t)kXEpi * kXCap * 32 * sizeof(uint2);
        CRW_CUDA_RET(cudaFuncSetAttribute(lp_filter_kernel<KT, KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_dev[dev] = smem;
    }
    const int items = p.v_end - p.v_begin;
    const int grid = items < max_ctas ? items : max_ctas;
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(plan.maps);
    lp_filter_kernel<KT, KL><<<grid, kXThreads, smem, st>>>(maps[0], maps[1], p);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

template <int SL, int WARPS, int SROWS>
static int launch_refine(const RParams& r, int max_ctas, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = (size_t)WARPS * SROWS * kRRowBytes + (size_t)SL * kRChunk * sizeof(int);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        CRW_CUDA_RET(cudaFuncSetAttribute(lp_refine_kernel<SL, WARPS, SROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const long long bufs = ((long long)r.R * (r.row_end - r.row_begin) + 32 / SL - 1) / (32 / SL);
    if (bufs <= 0) return CRW_OK;
    const int grid = (int)(bufs < max_ctas ? bufs : max_ctas);
    lp_refine_kernel<SL, WARPS, SROWS><<<grid, WARPS * 32, smem, st>>>(r);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// prep + plan.  feats [R,T,N,128] fp32 -> plan (tensor maps, schedule).  `sms` sizes the items so that one round fills the GPU.
int lp_x_prepare(const float* feats, int R, int T, int N, int C, int ctx, float radius, float temp, int k, int do_normalize, float* W,
                 int32_t* I, void* scratch, int sms, cudaStream_t st, void* plan_storage, size_t plan_bytes) {
    if (plan_bytes < sizeof(LpXPlan)) return CRW_ERR_WORKSPACE;
    LpXPlan* plan = reinterpret_cast<LpXPlan*>(plan_storage);
    if (C != 128 || N > 128 || N < 8 || k > lp_x_max_k()) return CRW_ERR_UNSUPPORTED;
    x_kt_kl(k, plan->kt, plan->kl);
    const size_t rows = (size_t)R * T * N;
    if (rows * 128 >= (1ull << 40)) return CRW_ERR_UNSUPPORTED;
    if ((uint64_t)T * N * N >= (1ull << 32)) return CRW_ERR_UNSUPPORTED;
    char* sp = reinterpret_cast<char*>(scratch);
    unsigned* stats = reinterpret_cast<unsigned*>(sp);
    sp += 256;
    __half* h = reinterpret_cast<__half*>(sp);
    sp += align_up(rows * C * sizeof(__half), 256);
    float* xn = nullptr;
    if (do_normalize) { xn = reinterpret_cast<float*>(sp); sp += align_up(rows * C * sizeof(float), 256); }
    int32_t* surv = reinterpret_cast<int32_t*>(sp);
    sp += align_up(rows * plan->kl * sizeof(int32_t), 256);
    int32_t* cnt = reinterpret_cast<int32_t*>(sp);

    CRW_CUDA_RET(cudaMemsetAsync(stats, 0, 256, st));
    {
        const int64_t blocks = ((int64_t)rows + 7) / 8;
        if (blocks > 0x7fffffffLL) return CRW_ERR_UNSUPPORTED;
        lp_prep_x_kernel<<<(unsigned)blocks, 256, 0, st>>>(feats, (int64_t)rows, do_normalize, xn, h, stats);
        CRW_LAUNCH_RET();
    }
    XParams& p = plan->p;
    p.R = R; p.T = T; p.N = N; p.ctx = ctx; p.k = k;
    p.rows_rg = T * N;
    const float rc_ = ceilf(radius);
    p.rb = (rc_ - 1.0f >= (float)N) ? N : (int)rc_ - 1;
    p.inv_temp = 1.0f / temp;
    p.magic_n = (unsigned)((1ull << 32) / (unsigned)N) + 1u;
    p.total_rows = (long long)rows;
    p.surv = surv; p.cnt = cnt; p.stats = stats;
    { const char* e = getenv("CRW_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
    // items: as many as fill whole rounds of the GPU, at most kXG * 128 rows each; the stream of one item must fit the 12-bit column
    const int max_rows = kXG * kXBM;
    long long items = ((long long)rows + max_rows - 1) / max_rows;
    if (sms > 0) items = (items + sms - 1) / sms * sms;
    int items_rg = (int)((items + R - 1) / R);
    if (items_rg < 1) items_rg = 1;
    int rpi = ceil_div(p.rows_rg, items_rg);
    if (rpi < 1) rpi = 1;
    // stream length check: frame-0 tiles + (rows of the item + ctx + 1 frames) in 64-row tiles must stay below 2^12 / 64 tiles
    while (true) {
        const int frames = rpi / N + 2 + ctx;
        const int ktiles = ceil_div(N, kXBN) + ceil_div(frames * N, kXBN) + 1;
        if (ktiles <= (1 << kXColBits) / kXBN) break;
        if (rpi <= N) return CRW_ERR_UNSUPPORTED;
        rpi = rpi / 2;
    }
    p.rows_per_item = rpi;
    p.items_per_rg = ceil_div(p.rows_rg, rpi);
    const int early = ceil_div(min((ctx + 2) * N, p.rows_rg), rpi);
    p.early_items_rg = early < p.items_per_rg ? early : p.items_per_rg;
    p.v_begin = p.v_end = 0;
    RParams& r = plan->r;
    r.xn = do_normalize ? xn : feats;
    r.surv = surv; r.cnt = cnt; r.W = W; r.I = I;
    r.R = R; r.T = T; r.N = N; r.ctx = ctx; r.rb = p.rb; r.k = k;
    r.rows_rg = p.rows_rg; r.row_begin = 0; r.row_end = 0;
    r.total_rows = p.total_rows; r.inv_temp = p.inv_temp; r.magic_n = p.magic_n; r.debug = p.debug;
    if (T < 2) return CRW_OK;
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(plan->maps);
    int rc = make_tmap_bf16_k64(&maps[0], h, (uint64_t)rows, 128, kXBM);      // 2-byte elements: the bf16 map type moves fp16 as well
    if (rc == CRW_OK) rc = make_tmap_bf16_k64(&maps[1], h, (uint64_t)rows, 128, kXBN);
    return rc;
}
int lp_x_total_slots(const void* plan) { const LpXPlan* pl = reinterpret_cast<const LpXPlan*>(plan); return pl->p.R * pl->p.items_per_rg; }
int lp_x_early_slots(const void* plan) { const LpXPlan* pl = reinterpret_cast<const LpXPlan*>(plan); return pl->p.R * pl->p.early_items_rg; }
int lp_x_early_rows(const void* plan) {
    const LpXPlan* pl = reinterpret_cast<const LpXPlan*>(plan);
    const int r = pl->p.early_items_rg * pl->p.rows_per_item;
    return r < pl->p.rows_rg ? r : pl->p.rows_rg;
}

// filter over schedule slots [v_begin, v_end), then refine over the rows those slots cover (early slots: rows [0, early_rows) of
// every radargram; the rest: [early_rows, rows_rg)); v ranges must be exactly the early part, the rest, or everything
int lp_x_launch(const void* plan_storage, int v_begin, int v_end, int max_ctas, cudaStream_t st) {
    const LpXPlan& plan = *reinterpret_cast<const LpXPlan*>(plan_storage);
    if (v_end <= v_begin) return CRW_OK;
    XParams p = plan.p;
    p.v_begin = v_begin;
    p.v_end = v_end;
    int rc;
    if (plan.kt == 10) rc = launch_filter<10, 16>(plan, p, max_ctas, st);
    else if (plan.kt == 16) rc = launch_filter<16, 24>(plan, p, max_ctas, st);
    else rc = launch_filter<24, 32>(plan, p, max_ctas, st);
    if (rc != CRW_OK) return rc;
    RParams r = plan.r;
    const int early_slots = p.R * p.early_items_rg, early_rows = lp_x_early_rows(plan_storage);
    r.row_begin = (v_begin >= early_slots) ? early_rows : 0;
    r.row_end = (v_end <= early_slots) ? early_rows : p.rows_rg;
    return plan.kl <= 16 ? launch_refine<16, 16, 24>(r, max_ctas, st) : launch_refine<32, 11, 33>(r, max_ctas, st);
}
size_t lp_x_plan_bytes() { return sizeof(LpXPlan); }
int lp_x_profile_read(unsigned long long* host_out, int reset) {
    if (host_out) CRW_CUDA_RET(cudaMemcpyFromSymbol(host_out, g_x_prof, sizeof(g_x_prof)));
    if (reset) {
        void* ptr = nullptr;
        CRW_CUDA_RET(cudaGetSymbolAddress(&ptr, g_x_prof));
        CRW_CUDA_RET(cudaMemset(ptr, 0, sizeof(g_x_prof)));
    }
    return CRW_OK;
}

}  // namespace crw

// profiling aid: copies the per-warp phase counters of the exact tensor path ([2 kernels][160 CTAs][18 warps][8] uint64) to host
// memory (synchronising) and clears them when `reset`
extern "C" int crw_debug_lp_x_profile(unsigned long long* host_out, int reset) { return crw::lp_x_profile_read(host_out, reset); }
