// labelprop_x.cu -- label-propagation affinity + top-k for sm_100a: tensor-core FILTER + exact fp32 REFINE.
//
// Replaces the same reference code as labelprop_f32.cu (batched_affinity, src/imported/maskedatt.py:151-175, with the radius
// mask of :232-245) and produces BIT-IDENTICAL W / I (hence masks and labels) to the pinned-order fp32 path and to
// oracle/crw_oracle.c, at tensor-core speed (precision = CRW_PREC_TC_EXACT):
//
//   lp_mu_x_kernel      the centre mu of each radargram's features from a strided sample of its rows (sub-megabyte read)
//   lp_prep_x_kernel    ONE pass over the features: F.normalize in the pinned order -> xn (fp32), the query plane fp16(256 xn), the
//                       key plane fp16(256 (xn - mu)) (q . (k - mu) ranks the keys like q . k, with a rounding error that scales
//                       with the spread of the features), per-call maxima of the norms and rounding residuals (they size the margin)
//   lp_filter_kernel    ONE tcgen05 pass (kind::f16, fp16 operands, fp32 accumulate in TMEM) over the dense (query tile x key
//                       tile) blocks.  The approximate dot a~ differs from the pinned fp32 chain dot by at most E (Cauchy-
//                       Schwarz on the rounding residuals + accumulation slack), so every candidate whose a~ is within 2E of
//                       the k-th best a~ survives: the true top-k is a subset of the survivors.  Four 128-row query tiles
//                       share every 64-row key stage (the L2 -> SM key stream is a quarter of one tile per stage).
//                         warp 16      TMA producer: ring of 16 KB stages (query half-tiles, then key tiles), SWIZZLE_128B
//                         warps 17-18  MMA issuers: query tiles copied smem -> TMEM (tcgen05.cp), TS-form MMAs, 4 accumulators
//                         warps 0-15   epilogue: thread = query row; branch-free "beats the bound and lies in the window/band"
//                                      test per value, survivors appended to a per-thread smem column; a warp-wide flush merges
//                                      them into a sorted register list of KL entries whose KT-th entry gives the bound.
//                                      Queries whose list overflowed go to the launch's overflow list in global memory
//   lp_refine_kernel    every warp on its own, chunks of up to 32 queries: phase 1 (warp per query) loads the fp32 rows of the
//                       survivors straight from L2 and forms the dots in the oracle's pinned order; phase 2 (lane per query)
//                       sorts them (logit desc, id asc), masked fill, pinned softmax, W / I stores.  Overflowed queries are
//                       rescanned in full, one per CTA at a time: a few dedicated CTAs start on the overflow list at once,
//                       the others join when their chunks are done -- slower, same result.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace crw {

// ------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------
constexpr float kXScale = 256.0f;            // fp16 plane holds 256 * xn: keeps small components out of the subnormal range

constexpr int kXStatSlots = 32;           // the per-call maxima are spread over 32 slots (fewer same-address atomics), reduced by the filter
constexpr int kXStatVals = 4;             // per slot: max |xn|^2, max |xn - hq/256|^2, max |xn - mu|^2, max |(xn - mu) - hk/256|^2
constexpr int kXStatBytes = kXStatSlots * kXStatVals * 4;
constexpr int kXCtrBytes = 64;             // overflow-list counters: [early | bulk launch][count, steal cursor]

// F.normalize of one row held as v[m] = channel lane + 32 m, in the pinned order (identical to l2_normalize_kernel /
// crw_oracle_l2_normalize)
__device__ __forceinline__ void x_normalize_row(float (&v)[4]) {
    float ss = 0.0f;
#pragma unroll
    for (int m = 0; m < 4; ++m) ss = __fmaf_rn(v[m], v[m], ss);
    ss = warp_sum_butterfly_rn(ss);
    const float d = fmaxf(__fsqrt_rn(ss), kNormEps);
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = __fdiv_rn(v[m], d);
}

// pass 0: the centre mu of a radargram's features, estimated from a strided SAMPLE of its rows (at most kXMuRows): ANY vector works as
// the centre -- q . (k - mu) ranks the keys of a query exactly like q . k, and the filter margin is computed from the residuals and
// norms actually measured against this mu in pass 1 -- the closer mu is to the mean, the shorter the survivor lists.  Sampling
// makes it a sub-megabyte read, so that the planes are written by ONE pass over the features.
constexpr int kXMuRows = 1024;      // one row per warp when the call holds few radargrams: the kernel is one HBM latency long
__global__ void __launch_bounds__(256) lp_mu_x_kernel(const float* __restrict__ x, int rows_rg, int do_normalize, float* __restrict__ musum) {
    __shared__ float s_col[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = blockIdx.y;
    const int nsample = min(rows_rg, kXMuRows);
    const int stride = rows_rg / nsample;                       // >= 1; sample i is row i * stride
    float col[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = blockIdx.x * 8 + warp; i < nsample; i += gridDim.x * 8) {
        const float* xr = x + ((int64_t)rg * rows_rg + (int64_t)i * stride) * 128;
        float v[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) v[m] = xr[lane + 32 * m];
        if (do_normalize) x_normalize_row(v);
#pragma unroll
        for (int m = 0; m < 4; ++m) col[m] += v[m];
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) s_col[warp][lane + 32 * m] = col[m];
    __syncthreads();
    if (threadIdx.x < 128) {
        float c = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) c += s_col[w][threadIdx.x];
        atomicAdd(&musum[rg * 128 + threadIdx.x], c * (1.0f / (float)nsample));
    }
}

// pass 1 (the only pass over all the features): F.normalize in the pinned order -> xn (fp32), QUERY plane hq = fp16(256 xn), KEY
// plane hk = fp16(256 (xn - mu)), and the per-call maxima that size the filter margin: |xn|^2, |xn - hq/256|^2, |xn - mu|^2,
// |(xn - mu) - hk/256|^2.  The rounding error of q . (k - mu) scales with |k - mu| instead of |k|: on near-collinear embeddings
// (any encoder at initialisation) the margin shrinks with the spread of the features and the survivor lists stay short.
// One warp per row, two rows per trip (eight independent loads in flight per lane); a CTA's rows belong to one radargram.
__global__ void __launch_bounds__(256) lp_prep_x_kernel(const float* __restrict__ x, int rows_rg, int do_normalize, float* __restrict__ xn,
                                                        __half* __restrict__ hq, __half* __restrict__ hk, const float* __restrict__ mu_all,
                                                        unsigned* __restrict__ stats) {
    __shared__ float s_st[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = blockIdx.y;
    float mu[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) mu[m] = mu_all[rg * 128 + lane + 32 * m];
    float mx[4] = {0.f, 0.f, 0.f, 0.f};
    const int step = gridDim.x * 16;
    for (int r0 = blockIdx.x * 16 + warp * 2; r0 < rows_rg; r0 += step) {
        const int nr = min(2, rows_rg - r0);
        float v[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float* xr = x + ((int64_t)rg * rows_rg + r0 + (u < nr ? u : 0)) * 128;
#pragma unroll
            for (int m = 0; m < 4; ++m) v[u][m] = xr[lane + 32 * m];
        }
        float acc[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (do_normalize) x_normalize_row(v[u]);
            const int64_t row = (int64_t)rg * rows_rg + r0 + u;
            float n2 = 0.0f, e2 = 0.0f, d2 = 0.0f, f2 = 0.0f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float val = v[u][m], dv = val - mu[m];
                const __half hv = __float2half_rn(val * kXScale), kv = __float2half_rn(dv * kXScale);
                const float res = val - __half2float(hv) * (1.0f / kXScale), rk = dv - __half2float(kv) * (1.0f / kXScale);
                n2 = fmaf(val, val, n2);
                e2 = fmaf(res, res, e2);
                d2 = fmaf(dv, dv, d2);
                f2 = fmaf(rk, rk, f2);
                if (u < nr) {
                    hq[row * 128 + lane + 32 * m] = hv;
                    hk[row * 128 + lane + 32 * m] = kv;
                    if (xn) xn[row * 128 + lane + 32 * m] = val;
                }
            }
            acc[u][0] = n2; acc[u][1] = e2; acc[u][2] = d2; acc[u][3] = f2;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[u][j] += __shfl_xor_sync(0xffffffffu, acc[u][j], off);
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[j] = fmaxf(mx[j], fmaxf(acc[0][j], nr > 1 ? acc[1][j] : 0.0f));
    }
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_st[warp][j] = mx[j];
    __syncthreads();
    if (threadIdx.x < 4) {          // non-negative floats order like their bit patterns
        float a = 0.0f;
        for (int w = 0; w < 8; ++w) a = fmaxf(a, s_st[w][threadIdx.x]);
        atomicMax(&stats[kXStatVals * (blockIdx.x % kXStatSlots) + threadIdx.x], __float_as_uint(a));
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
constexpr int kXThreads = (kXEpi + 4) * 32;         // 16 epilogue warps + one warp group: TMA producer, two MMA issuers, one idle
constexpr int kXEpiRegs = 112, kXAuxRegs = 32;      // setmaxnreg: the epilogue warp groups take the registers the fifth one gives up
constexpr int kXColBits = 12;                // packed key: value << 12 | stream column (key tile * 64 + column)
constexpr float kXFixHalf = 524280.0f;       // 2^19 - 8: fix = a * scale + kXFixHalf lies in [0, 2^20) for |a| <= the call's bound
constexpr float kXFixMagic = 12582912.0f;    // 1.5 * 2^23: float -> integer in the low mantissa bits

struct XParams {
    int R, T, N, ctx, rb, k;
    int n_min;                               // first query frame whose results are wanted (1; ctx + 1 for a streamed sub-sequence)
    int rows_rg;                             // T * N
    int rows_per_item, items_per_rg, early_items_rg;
    int early_rpi, early_rows;               // the first early_items_rg items of a radargram are small (early_rpi rows): rows [0, early_rows)
    int v_begin, v_end;                      // schedule slots walked by this launch
    unsigned magic_n;
    int debug;
    long long total_rows;
    int32_t* surv;                           // [total_rows][KL] radargram-relative key rows of the survivors (sorted prefix of the list)
    int32_t* cnt;                            // [total_rows]     number of survivors, bit 30 = list overflowed (rescan in full)
    int32_t* ovf_list;                       // global rows of the overflowed queries of this launch, in arrival order
    int* ovf_ctr;                            // [0] number of entries of ovf_list, [1] the refine kernel's steal cursor
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
    if (it < p.early_items_rg) {
        t.ra = it * p.early_rpi;
        t.rb = min(t.ra + p.early_rpi, p.early_rows);
    } else {
        t.ra = p.early_rows + (it - p.early_items_rg) * p.rows_per_item;
        t.rb = min(t.ra + p.rows_per_item, p.rows_rg);
    }
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
__device__ __forceinline__ float x_threshold(uint32_t theta, int m_fix, float inv_scale) {
    if (theta == 0u) return -INFINITY;
    const int X = (int)(theta >> kXColBits) - m_fix - 2;             // fixed-point steps of slack for the rounding of fix() and of this product
    return ((float)max(X, 0) - kXFixHalf) * inv_scale;
}

// 16 columns of a key tile: branch-free append of the values that beat the bound and are valid (window x band).
// Per value: validity bit -> predicate, compare, column id, predicated 8-byte store (raw value, column), predicated slot advance.
// Written as PTX so that the per-value work stays at these five instructions (the fixed-point packing happens at flush time).
template <int I, int END, int OFF>
struct XScan {
    static __device__ __forceinline__ void run(const float (&val)[64], uint32_t vbits, uint32_t colbase, float thr, uint32_t& ptr) {
        asm volatile(
            "{\n\t.reg .pred p, v;\n\t.reg .b32 t, c;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 v, t, 0;\n\t"
            "setp.ge.and.f32 p, %3, %4, v;\n\t"
            "or.b32 c, %5, %6;\n\t"
            "@p st.shared.v2.b32 [%0], {%7, c};\n\t"
            "@p add.u32 %0, %0, 256;\n\t}"
            : "+r"(ptr)
            : "r"(vbits), "n"(1u << I), "f"(val[OFF + I]), "f"(thr), "r"(colbase), "n"(I), "r"(__float_as_uint(val[OFF + I]))
            : "memory");
        XScan<I + 1, END, OFF>::run(val, vbits, colbase, thr, ptr);
    }
};
template <int END, int OFF>
struct XScan<END, END, OFF> {
    static __device__ __forceinline__ void run(const float (&)[64], uint32_t, uint32_t, float, uint32_t&) {}
};
// 16 columns I0 .. I0 + 15 of half H (columns 32 H ..) of a key tile; vbits = the validity bits of that half
template <int H, int I0>
__device__ __forceinline__ void x_scan16(const float (&val)[64], uint32_t vbits, uint32_t colbase, float thr, uint32_t& ptr) {
    XScan<I0, I0 + 16, 32 * H>::run(val, vbits, colbase, thr, ptr);
}

template <int KT, int KL>
__global__ void __launch_bounds__(kXThreads, 1)
lp_filter_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap, XParams p) {
    constexpr int kProducerWarp = kXEpi, kMmaWarp = kXEpi + 1;      // issuers: warps kMmaWarp (tiles 0, 1) and kMmaWarp + 1 (tiles 2, 3)
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
        for (int s = 0; s < kXStages; ++s) { tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 2); }   // empty: one commit per issuer
        for (int g = 0; g < kXG; ++g) { tc::mbar_init(&acc_full[g], 1); tc::mbar_init(&acc_empty[g], 4); }
        tc::fence_barrier_init();
    }
    if (warp == kProducerWarp && lane == 0) { tc::prefetch_tmap(&qmap); tc::prefetch_tmap(&kmap); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;      // columns [0,256): four query tiles (64 each); [256,512): four accumulators
    // register budget (setmaxnreg, first thing in every role): the epilogue keeps a whole 64-column key tile in registers so
    // that it can free the accumulator at once; the auxiliary warp group gives up what it does not need
    if (warp >= kXEpi) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kXAuxRegs));
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
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
        // ================= MMA issuers: warp kMmaWarp serves query tiles 0 and 1, warp kMmaWarp + 1 tiles 2 and 3 =================
        // Both walk every ring slot (waiting until it is full, so that neither runs ahead of the producer) and commit to its
        // `empty` barrier (count 2): the commit covers whatever the issuer queued for that slot -- the tcgen05.cp of one of its
        // query tiles, the MMAs of the tiles that look at the key tile, or nothing.
        const int iw = warp - kMmaWarp;
        const bool leader = tc::elect_one();
        const uint64_t kdesc0 = tc::umma_smem_desc_k128(tc::smem_u32(sK));
        uint32_t scnt = 0, ucnt[2] = {0u, 0u};
        for (int v = p.v_begin + blockIdx.x; v < p.v_end; v += gridDim.x) {
            const XItem t = x_item(p, v);
            if (t.n_kt == 0) continue;
            XTile qt[2];
#pragma unroll
            for (int gl = 0; gl < 2; ++gl) qt[gl] = x_tile(p, t, 2 * iw + gl);
            // query tiles: ring slots -> TMEM (tcgen05.cp executes in order with the MMAs this thread issues before and after it)
            for (int g = 0; g < kXG; ++g) {
                if (!x_tile(p, t, g).active) continue;
                for (int kb = 0; kb < 2; ++kb, ++scnt) {
                    const int s = scnt % kXStages;
                    tc::mbar_wait(&k_full[s], (scnt / kXStages) & 1);
                    tc::tc_fence_after();
                    if (leader) {
                        if ((g >> 1) == iw) {
                            const uint64_t sd = kdesc0 + (uint64_t)((s * kXStageBytes) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc::tmem_cp_128x256b(tmem_base + (uint32_t)(g * 64 + kb * 32 + ks * 8), sd + (uint64_t)((ks * 32) >> 4));
                        }
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
                for (int gl = 0; gl < 2; ++gl) {
                    if (!x_needs(p, qt[gl], row0, nrows)) continue;
                    const int g = 2 * iw + gl;
                    tc::mbar_wait(&acc_empty[g], ((ucnt[gl]) & 1) ^ 1);
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
                    ++ucnt[gl];
                }
                if (leader) tc::umma_commit(&k_empty[s]);
                __syncwarp();
            }
        }
    }
      // (the fourth warp of the auxiliary warp group is idle)
    } else {
        // ================= epilogue: thread = query row =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kXEpiRegs));
        const int g = warp >> 2, quarter = warp & 3;
        const int rb = p.rb, ctx = p.ctx;
        const uint32_t app = tc::smem_u32(sApp + (size_t)warp * kXCap * 32) + lane * 8;
        // margin from the call's maxima: a~ = tensor(hq . hk) / 65536 approximates q . (k - mu); with eq, ed the largest rounding
        // residual norms of the query / centred key planes and nq, nd the largest norms,
        //   |a~ - (chain(q, k) - q . mu)| <= E = eq nd + nq ed + eq ed + (tensor accumulation + chain rounding + the fp32 subtraction)
        // so every candidate within 2 E of the KT-th best a~ survives (the true top-k is among them).
        float nq2 = 0.0f, eq2 = 0.0f, nd2 = 0.0f, ed2 = 0.0f;
        for (int i = 0; i < kXStatSlots; ++i) {
            nq2 = fmaxf(nq2, __uint_as_float(p.stats[kXStatVals * i]));
            eq2 = fmaxf(eq2, __uint_as_float(p.stats[kXStatVals * i + 1]));
            nd2 = fmaxf(nd2, __uint_as_float(p.stats[kXStatVals * i + 2]));
            ed2 = fmaxf(ed2, __uint_as_float(p.stats[kXStatVals * i + 3]));
        }
        const float nq = sqrtf(nq2) * 1.00001f, eq = sqrtf(eq2) * 1.0001f + 1e-9f;
        const float nd = sqrtf(nd2) * 1.00001f + 1e-9f, ed = sqrtf(ed2) * 1.0001f + 1e-9f;
        const float E = (eq * nd + nq * ed + eq * ed + 1.0e-5f * nq * nq + 2.0e-5f * nq * nd) * 1.02f;
        // fixed point: a (accumulator units, 65536 x dot) lies within +-65536 A, A = 1.05 nq nd; 19 bits cover A
        const float A = 1.05f * nq * nd + 1e-6f;
        const float fix_scale = kXFixHalf / (A * 65536.0f), inv_scale = 1.0f / fix_scale;
        const int m_fix = (int)ceilf(2.0f * E * 65536.0f * fix_scale) + 2;
        const bool range_bad = !(nq2 <= 1.002f) || m_fix > 200000;   // un-normalised input / degenerate scale: rescan everything
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
            const bool qvalid = qt.active && in_item && (n >= p.n_min) && (n < p.T);
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
                    const float tf = __fmaf_rn(__uint_as_float(vb), fix_scale, kXFixHalf + kXFixMagic);
                    const uint32_t P = (__float_as_uint(tf) << kXColBits) | col;
                    x_list_insert<KL>(L, (i < cnt) ? P : 0u);
                }
                ptr = app;
                thr = x_threshold(L[KT - 1], m_fix, inv_scale);
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
                float val[64];
                if (!(p.debug & 2)) {
                    // both 32-column loads in flight before the one wait; once the values sit in registers the accumulator goes
                    // straight back to the issuer, whose next MMAs then run under the whole scan of this key tile
                    tc::tmem_ld_32x32b_x32(taddr, reinterpret_cast<float(&)[32]>(val[0]));
                    if (nrows > 32) tc::tmem_ld_32x32b_x32(taddr + 32u, reinterpret_cast<float(&)[32]>(val[32]));
                    tc::tmem_ld_wait();
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[g]);
                if (!(p.debug & 2)) {
                    x_scan16<0, 0>(val, (uint32_t)vm, (uint32_t)(kt * 64), thr, ptr);
                    maybe_flush();
                    x_scan16<0, 16>(val, (uint32_t)vm, (uint32_t)(kt * 64), thr, ptr);
                    maybe_flush();
                    if (nrows > 32) {
                        x_scan16<1, 0>(val, (uint32_t)(vm >> 32), (uint32_t)(kt * 64 + 32), thr, ptr);
                        maybe_flush();
                        x_scan16<1, 16>(val, (uint32_t)(vm >> 32), (uint32_t)(kt * 64 + 32), thr, ptr);
                        maybe_flush();
                    }
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
                    // the list is sorted: the survivors are a prefix of it.  All KL key rows are written (16-byte stores, row-major
                    // [row][KL]); the count says how many of them matter
                    int krs[KL];
#pragma unroll
                    for (int s = 0; s < KL; ++s) {
                        const int col = (int)(L[s] & ((1u << kXColBits) - 1u));
                        const int kt2 = col >> 6, i = col & 63;
                        krs[s] = (kt2 < t.has_f0) ? kt2 * kXBN + i : t.f_lo * N + (kt2 - t.has_f0) * kXBN + i;
                        c += (L[s] != 0u && L[s] >= bound) ? 1 : 0;
                    }
                    int4* dst = reinterpret_cast<int4*>(p.surv + grow * KL);
#pragma unroll
                    for (int s = 0; s < KL; s += 4) dst[s >> 2] = make_int4(krs[s], krs[s + 1], krs[s + 2], krs[s + 3]);
                }
                p.cnt[grow] = qvalid ? (c | (ovf ? (1 << 30) : 0)) : 0;
                if (qvalid && ovf) p.ovf_list[atomicAdd(p.ovf_ctr, 1)] = (int32_t)grow;
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
constexpr int kRWarps = 8;                   // warps per CTA, one query at a time each
constexpr int kRThreads = kRWarps * 32;
constexpr int kRRescCtas = 16;                // CTAs of the refine grid that only serve the overflow list

struct RParams {
    const float* xn;         // [total_rows, 128] normalised features
    const int32_t* surv;     // [total_rows][KL] key rows of the survivors, best approximate value first
    const int32_t* cnt;      // [total_rows]
    float* W;                // [R, T, k, N]
    int32_t* I;
    int R, T, N, ctx, rb, k;
    int n_min;               // first query frame whose results are wanted
    int rows_rg, row_begin, row_end;   // radargram-relative row range handled by this launch (every radargram)
    long long total_rows;
    float inv_temp;
    unsigned magic_n;
    int chunk_q;             // queries per warp chunk of the refine kernel (<= 32)
    int resc_ctas;           // the first resc_ctas CTAs of the grid take no chunks: they start on the overflow list at once
    const int32_t* ovf_list; // global rows of the overflowed queries of this launch (written by the filter)
    int* ovf_ctr;            // [0] their number, [1] steal cursor
    int debug;               // timing aids (CRW_TC_DEBUG): 16 = no W / I stores, 32 = no row loads, 64 = no dot products (results invalid)
};

// candidate id (slot in the trimmed key set * N + node) of key row kr for query frame n
__device__ __forceinline__ int x_cand_id(int kr, int n, int N, int ctx, unsigned magic_n) {
    const int kf = (int)__umulhi((unsigned)kr, magic_n), j = kr - kf * N;
    const int slot = (n > ctx + 1 && kf != 0) ? kf - (n - ctx) + 1 : kf;
    return slot * N + j;
}

// lanes [0, k) hold a sorted list (logit desc, id asc), one entry per lane; insert (c, ci) keeping that order.  Candidates may
// arrive in any order (full comparator).
__device__ __forceinline__ void x_list_insert_warp(float& v, int& id, float c, int ci, int k, unsigned kmask, int lane) {
    const int pos = __popc(__ballot_sync(0xffffffffu, v > c || (v == c && id < ci)) & kmask);
    if (pos < k) {
        const float vup = __shfl_up_sync(0xffffffffu, v, 1);
        const int iup = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane > pos) { v = vup; id = iup; }
        else if (lane == pos) { v = c; id = ci; }
        if (lane >= k) v = -INFINITY;
    }
}

// Exact scan of the in-band candidates number c = round_begin * 16 + 16 j ... of one query by one warp (its survivor list
// overflowed: many candidates tie within the filter margin): rounds of 16 candidates, their rows loaded together, one warp dot
// each; lane i keeps the i-th best seen by this warp.  round_stride > 1: the rounds are dealt to several warps.
__device__ __forceinline__ void x_full_scan(const RParams& p, const float* xr, int n, int q, int k, int round_begin, int round_stride,
                                            float& v_out, int& id_out) {
    const int lane = threadIdx.x & 31;
    const int N = p.N, rb = p.rb;
    const int F = n_key_frames(n, p.ctx);
    const unsigned kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    float v = -INFINITY;
    int id = 0x7fffffff;
    const int jlo = max(0, q - rb), jhi = min(N - 1, q + rb), bw = jhi - jlo + 1;
    const unsigned magic_bw = (unsigned)((1ull << 32) / (unsigned)bw) + 1u;      // cc / bw for cc * bw < 2^32
    const float4 qv = __ldg(reinterpret_cast<const float4*>(xr + ((size_t)n * N + q) * 128) + lane);
    const int total = F * bw;                          // in-band candidates, ascending id = (frame slot, node)
    for (int c0 = round_begin * 16; c0 < total; c0 += round_stride * 16) {
        const int nrow = min(16, total - c0);
        // sixteen unconditional loads, issued back to back (indices past the end repeat the last candidate: ignored below)
        float4 kv[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int cc = min(c0 + r, total - 1), f = (int)__umulhi((unsigned)cc, magic_bw), jj = cc - f * bw;
            kv[r] = __ldg(reinterpret_cast<const float4*>(xr + ((size_t)key_frame(n, p.ctx, f) * N + jlo + jj) * 128) + lane);
        }
        float part[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) part[r] = x_partial4(kv[r], qv);
        const float mine = __fmul_rn(x_warp_dot16(part, lane), p.inv_temp);           // candidate c0 + (lane >> 1)
        const int cc_l = c0 + (lane >> 1), f_l = (int)__umulhi((unsigned)min(cc_l, total - 1), magic_bw);
        const int id_l = f_l * N + jlo + (cc_l - f_l * bw);
        // only the candidates that beat the k-th best so far are inserted (ascending lane = ascending id); after the first
        // rounds that is a handful per round
        const float vk = __shfl_sync(0xffffffffu, v, k - 1);
        const int idk = __shfl_sync(0xffffffffu, id, k - 1);
        unsigned m = __ballot_sync(0xffffffffu, !(lane & 1) && (lane >> 1) < nrow && (mine > vk || (mine == vk && id_l < idk)));
        while (m) {                                                                    // warp-uniform
            const int src = __ffs(m) - 1;
            m &= m - 1;
            x_list_insert_warp(v, id, __shfl_sync(0xffffffffu, mine, src), __shfl_sync(0xffffffffu, id_l, src), k, kmask, lane);
        }
    }
    v_out = v;
    id_out = id;
}

// lane sl holds the winner of rank sl, (logit desc, id asc), `live` of them real: masked fill, pinned softmax (sequential sum
// in rank order), W / I stores.  es: k floats of scratch of this warp.
__device__ __forceinline__ void x_finish_query(const RParams& p, float v, int id, int live, int rg, int n, int q, int k, int sl, float* es) {
    const int N = p.N, rb = p.rb;
    if (sl >= live && sl < k) {
        // fewer than k in-band candidates: out-of-band ones share one logit and come in ascending id order
        const int lo_q = max(0, q - rb), w_q = min(N - 1, q + rb) - lo_q + 1, nob = max(N - w_q, 1);
        const int t = sl - live, f = t / nob, r = t - f * nob;
        const int jj = (r < lo_q) ? r : r + w_q;
        v = __fmul_rn(kMaskBias, p.inv_temp);
        id = f * N + jj;
    }
    const float v0 = __shfl_sync(0xffffffffu, v, 0);
    const bool w = sl < k;
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

// ---- thread-per-query finish: sort the survivors of one query (logit desc, id asc) and write its softmax weights ----
// 64-bit sort key: orderable logit bits in the upper word (larger = better), ~id in the lower word (smaller id = better); ids of
// a query are distinct, so keys are distinct and the order is the oracle's total order.  (Dot products come out of fmaf chains
// that start from +0 and butterfly sums: they are never -0, and neither is dot * (1 / temp).)
__device__ __forceinline__ unsigned long long x_sort_key(float lg, int id) {
    const uint32_t u = __float_as_uint(lg + 0.0f);
    const uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)o << 32) | (unsigned long long)(0xffffffffu - (uint32_t)id);
}
__device__ __forceinline__ float x_key_logit(unsigned long long key) {
    const uint32_t o = (uint32_t)(key >> 32);
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ int x_key_id(unsigned long long key) { return (int)(0xffffffffu - (uint32_t)key); }
// bitonic network over NS (power of two) register-resident keys, descending; every index is a compile-time constant
template <int NS>
__device__ __forceinline__ void x_sort_desc(unsigned long long (&key)[NS]) {
#pragma unroll
    for (int kk = 2; kk <= NS; kk <<= 1)
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool desc = (i & kk) == 0;
                    const unsigned long long a = key[i], b2 = key[l];
                    const bool sw = desc ? (a < b2) : (a > b2);
                    key[i] = sw ? b2 : a;
                    key[l] = sw ? a : b2;
                }
            }
}

// KL = survivor slots per query (16, 24 or 32).  Every WARP works on its own: it takes chunks of up to 32 consecutive queries and
// runs two phases on each, with no CTA-wide barrier between them (warps of an SM are in different phases at any time, so the
// latency-bound phase 2 of one runs under the loads of the others):
//   phase 1, the warp on one query at a time: the fp32 rows of the survivors come straight from L2 into registers (one coalesced
//            512-byte row per load instruction, all of a query's loads in flight together), every dot product is one warp dot in
//            the oracle's order (sixteen per reduce-scatter); the dots go to the warp's slab of shared memory.
//   phase 2, one LANE per query: exact order of the survivors (sorting network on registers), masked fill, pinned softmax
//            (sequential sum in rank order), W / I stores (consecutive lanes = consecutive nodes) -- no shuffles.
// Queries whose survivor list overflowed are queued and rescanned in full by the whole CTA at the end (rare).
template <int KL, int RB, int MINB>
__global__ void __launch_bounds__(kRThreads, MINB) lp_refine_kernel(RParams p) {
    constexpr int NS = (KL <= 16) ? 16 : 32;
    __shared__ float s_lg[kRWarps][KL][33];
    __shared__ float es_all[kRWarps][32];
    __shared__ float sc_v[kRWarps][32];
    __shared__ int sc_i[kRWarps][32];
    __shared__ int s_steal;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, k = p.k;
    const unsigned kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    const int rows_launch = p.row_end - p.row_begin;
    const int total_q = p.R * rows_launch;          // (32-bit arithmetic: the host checks R * rows < 2^31)
    const int qw = p.chunk_q;                       // queries per warp chunk (<= 32)
    const int nchunks = (total_q + qw - 1) / qw;
    const bool prof = (p.debug & 8) != 0;
    const int reg_ctas = (int)gridDim.x - p.resc_ctas;
    for (int ch = ((int)blockIdx.x - p.resc_ctas) * kRWarps + warp; ch < nchunks && (int)blockIdx.x >= p.resc_ctas; ch += reg_ctas * kRWarps) {
        const long long c_0 = prof ? clock64() : 0;
        const int r_lo = ch * qw, nq = min(total_q, r_lo + qw) - r_lo;
        // lane j owns query r_lo + j in phase 2; its geometry and survivor count are fetched once (coalesced)
        const int fj = r_lo + min(lane, nq - 1);
        const int rg_j = fj / rows_launch, row_j = p.row_begin + fj - rg_j * rows_launch;
        const int n_j = (int)__umulhi((unsigned)row_j, p.magic_n), q_j = row_j - n_j * N;
        const size_t grow_j = (size_t)rg_j * p.rows_rg + row_j;
        // count word: bit 31 = no query here (frame 0 / beyond the sequence / beyond the chunk), bit 30 = overflowed list
        unsigned cword_j = 0x80000000u;
        if (lane < nq && n_j >= p.n_min && n_j < p.T) cword_j = (unsigned)__ldg(p.cnt + grow_j);
        // ---------------- phase 1 ----------------
        // the survivor list of the NEXT query is fetched while the current one is worked on
        const size_t grow0 = (size_t)__shfl_sync(0xffffffffu, rg_j, 0) * p.rows_rg + __shfl_sync(0xffffffffu, row_j, 0);
        int kr_n = (lane < KL) ? __ldg(p.surv + grow0 * KL + lane) : 0;
        for (int i = 0; i < nq; ++i) {
            const unsigned cword = __shfl_sync(0xffffffffu, cword_j, i);
            const int rg = __shfl_sync(0xffffffffu, rg_j, i), row = __shfl_sync(0xffffffffu, row_j, i);
            const int kr = kr_n;
            {
                const int i1 = min(i + 1, nq - 1);
                const size_t grow1 = (size_t)__shfl_sync(0xffffffffu, rg_j, i1) * p.rows_rg + __shfl_sync(0xffffffffu, row_j, i1);
                kr_n = (lane < KL) ? __ldg(p.surv + grow1 * KL + lane) : 0;           // (slots beyond the count hold stale rows: unused)
            }
            if (cword & 0x80000000u) continue;                                  // warp-uniform
            if (cword & 0x40000000u) continue;                                  // overflowed list: on the launch's overflow list
            const int c = (int)(cword & 0xffffu);
            const float* xr = p.xn + (size_t)rg * p.rows_rg * 128;
            const float4 qv = __ldg(reinterpret_cast<const float4*>(xr + (size_t)row * 128) + lane);
#pragma unroll
            for (int sb = 0; sb < KL; sb += 16) {                            // (c <= KL: the slots beyond it are never touched)
                if (sb >= c) break;                                              // warp-uniform
                // the row loads of a batch (RB = 16: all of them, RB = 8: two batches) are unconditional and issued back to back
                // (slots past the count repeat the last survivor's row: same address, merged in L1, and their dots are
                // ignored), so that they are in flight together (measured: loads predicated on `slot < count` instead -- no duplicate
                // L1 traffic -- are issued one by one by the compiler: 73 us instead of 43)
                float part[16];
#pragma unroll
                for (int b0 = 0; b0 < 16; b0 += RB) {
                    float4 kv[RB];
#pragma unroll
                    for (int s2 = 0; s2 < RB; ++s2) {
                        int kr_s = __shfl_sync(0xffffffffu, kr, min(sb + b0 + s2, c - 1));
                        if (p.debug & 32) kr_s = row;
                        kv[s2] = __ldg(reinterpret_cast<const float4*>(xr + (size_t)kr_s * 128) + lane);
                    }
#pragma unroll
                    for (int s2 = 0; s2 < RB; ++s2) part[b0 + s2] = x_partial4(kv[s2], qv);
                }
                const float mine = x_warp_dot16(part, lane);                     // dot of slot sb + (lane >> 1)
                if (!(lane & 1) && sb + (lane >> 1) < KL) s_lg[warp][sb + (lane >> 1)][i] = mine;
            }
        }
        __syncwarp();
        const long long c_1 = prof ? clock64() : 0;
        // ---------------- phase 2 ----------------
        if (!(cword_j & 0xc0000000u)) {
            const int c = (int)(cword_j & 0xffffu), n = n_j, q = q_j, rg = rg_j;
            unsigned long long key[NS];
            const int4* sv = reinterpret_cast<const int4*>(p.surv + grow_j * KL);
#pragma unroll
            for (int s4 = 0; s4 < NS; s4 += 4) {
                int krs[4] = {0, 0, 0, 0};
                if (s4 < KL && s4 < c) { const int4 t4 = __ldg(sv + (s4 >> 2)); krs[0] = t4.x; krs[1] = t4.y; krs[2] = t4.z; krs[3] = t4.w; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int s2 = s4 + u;
                    unsigned long long kk = 0ull;
                    if (s2 < KL && s2 < c)
                        kk = x_sort_key(__fmul_rn(s_lg[warp][s2 < KL ? s2 : 0][lane], p.inv_temp), x_cand_id(krs[u], n, N, p.ctx, p.magic_n));
                    key[s2] = kk;
                }
            }
            x_sort_desc<NS>(key);
            const int live = min(c, k);
            // fewer than k in-band candidates: out-of-band ones share one logit and come in ascending id order
            const int lo_q = max(0, q - p.rb), w_q = min(N - 1, q + p.rb) - lo_q + 1, nob = max(N - w_q, 1);
            const float vfill = __fmul_rn(kMaskBias, p.inv_temp);
            const float v0 = (live > 0) ? x_key_logit(key[0]) : vfill;
            float e[KL];
            float ssum = 0.0f;
#pragma unroll
            for (int s2 = 0; s2 < KL; ++s2) {
                if (s2 < k) {
                    const float v = (s2 < live) ? x_key_logit(key[s2]) : vfill;
                    e[s2] = pinned_expf(__fsub_rn(v, v0));
                    ssum = (s2 == 0) ? e[0] : __fadd_rn(ssum, e[s2]);
                } else {
                    e[s2] = 0.0f;
                }
            }
            if (!(p.debug & 16)) {
                const size_t o0 = ((size_t)(rg * p.T + n) * k) * N + q;
#pragma unroll
                for (int s2 = 0; s2 < KL; ++s2) {
                    if (s2 < k) {
                        int id;
                        if (s2 < live) {
                            id = x_key_id(key[s2]);
                        } else {
                            const int t = s2 - live, fo = t / nob, r = t - fo * nob;
                            id = fo * N + ((r < lo_q) ? r : r + w_q);
                        }
                        p.W[o0 + (size_t)s2 * N] = __fdiv_rn(e[s2], ssum);
                        p.I[o0 + (size_t)s2 * N] = id;
                    }
                }
            }
        }
        __syncwarp();
        if (prof) { const long long c_2 = clock64(); XPROF(1, 1, c_1 - c_0); XPROF(1, 3, c_2 - c_1); }
    }
    __syncthreads();
    // ---------------- overflowed queries: rescanned in full, one at a time by the whole CTA (every warp scans its share of the
    // candidate rounds, warp 0 merges the partial lists).  The list was completed by the filter launch; entries are handed out
    // through a cursor: the first resc_ctas CTAs start here at once (a few overflows are absorbed while the others refine),
    // everyone else joins when its chunks are done (degenerate inputs overflow everywhere). ----------------
    const int novf = *reinterpret_cast<const volatile int*>(p.ovf_ctr);
    if (novf == 0) return;
    while (true) {
        __syncthreads();
        if (tid == 0) s_steal = atomicAdd(p.ovf_ctr + 1, 1);
        __syncthreads();
        const int e2 = s_steal;
        if (e2 >= novf) break;
        const int grow = p.ovf_list[e2];
        const int rg = grow / p.rows_rg, row = grow - rg * p.rows_rg;
        const int n = row / N, q = row - n * N;
        float pv;
        int pi;
        x_full_scan(p, p.xn + (size_t)rg * p.rows_rg * 128, n, q, k, warp, kRWarps, pv, pi);
        sc_v[warp][lane] = pv;
        sc_i[warp][lane] = pi;
        __syncthreads();
        if (warp == 0) {
            float v = -INFINITY;
            int id = 0x7fffffff;
            for (int w2 = 0; w2 < kRWarps; ++w2)
                for (int s2 = 0; s2 < k; ++s2) {
                    const float c = sc_v[w2][s2];
                    const int ci = sc_i[w2][s2];
                    if (!(c > -INFINITY)) break;                              // (warp-uniform) lists are sorted: the rest is empty
                    const float vk = __shfl_sync(0xffffffffu, v, k - 1);
                    const int idk = __shfl_sync(0xffffffffu, id, k - 1);
                    if (!(c > vk || (c == vk && ci < idk))) break;           // ... or cannot enter the merged list any more
                    x_list_insert_warp(v, id, c, ci, k, kmask, lane);
                }
            const int live = __popc(__ballot_sync(0xffffffffu, v > -INFINITY) & kmask);
            x_finish_query(p, v, id, live, rg, n, q, k, lane, es_all[0]);
        }
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

// scratch layout of the exact tensor path: stats + column sums | fp16 query plane | fp16 key plane | xn (when normalising) | survivors | counts
size_t lp_x_scratch_bytes(int R, int T, int N, int C, int k, int do_normalize) {
    int kt, kl;
    x_kt_kl(k, kt, kl);
    const size_t rows = (size_t)R * T * N;
    size_t b = align_up((size_t)kXStatBytes + kXCtrBytes + (size_t)R * 128 * sizeof(float), 256);
    b += 2 * align_up(rows * C * sizeof(__half), 256);
    b += align_up(rows * sizeof(int32_t), 256);          // overflow lists
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

template <int KL>
static int launch_refine(const RParams& r_in, int max_ctas, cudaStream_t st) {
    const RParams& r0 = r_in;
    const long long total_q = (long long)r0.R * (r0.row_end - r0.row_begin);
    if (total_q <= 0) return CRW_OK;
    if (total_q >= (1ll << 31) - 65536) return CRW_ERR_UNSUPPORTED;
    // two CTAs of 8 warps per SM; a warp chunk is a contiguous range of at most 32 queries, sized so that one round of
    // chunks fills every warp slot when the launch is small
    const long long slots = 2LL * max_ctas;
    const long long resc = (slots >= 8 * kRRescCtas) ? kRRescCtas : 1;      // CTAs that only serve the overflow list
    const long long wslots = (slots - resc) * kRWarps;
    long long per = (total_q + wslots - 1) / wslots;
    if (per > 32) per = 32;
    if (per < 8) per = 8;
    const long long nchunks = (total_q + per - 1) / per;
    long long ctas = (nchunks + kRWarps - 1) / kRWarps;
    if (ctas > slots - resc) ctas = slots - resc;
    ctas += resc;
    RParams r = r_in;
    r.chunk_q = (int)per;
    r.resc_ctas = (int)resc;
    // (RB = 8 with three CTAs per SM was measured: 80 vs 76 us, the 80-register budget spills)
    lp_refine_kernel<KL, 16, 2><<<(unsigned)ctas, kRThreads, 0, st>>>(r);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// prep + plan.  feats [R,T,N,128] fp32 -> plan (tensor maps, schedule).  `sms` sizes the items so that one round fills the GPU.
int lp_x_prepare(const float* feats, int R, int T, int N, int C, int ctx, float radius, float temp, int k, int do_normalize, float* W,
                 int32_t* I, void* scratch, int sms, cudaStream_t st, void* plan_storage, size_t plan_bytes, int n_min) {
    if (plan_bytes < sizeof(LpXPlan)) return CRW_ERR_WORKSPACE;
    LpXPlan* plan = reinterpret_cast<LpXPlan*>(plan_storage);
    if (C != 128 || N > 128 || N < 8 || k > lp_x_max_k()) return CRW_ERR_UNSUPPORTED;
    x_kt_kl(k, plan->kt, plan->kl);
    const size_t rows = (size_t)R * T * N;
    if (rows * 128 >= (1ull << 40)) return CRW_ERR_UNSUPPORTED;
    if ((uint64_t)T * N * N >= (1ull << 32)) return CRW_ERR_UNSUPPORTED;
    char* sp = reinterpret_cast<char*>(scratch);
    unsigned* stats = reinterpret_cast<unsigned*>(sp);
    int* ovf_ctr = reinterpret_cast<int*>(sp + kXStatBytes);
    float* musum = reinterpret_cast<float*>(sp + kXStatBytes + kXCtrBytes);
    const size_t head = align_up((size_t)kXStatBytes + kXCtrBytes + (size_t)R * 128 * sizeof(float), 256);
    sp += head;
    __half* hq = reinterpret_cast<__half*>(sp);
    sp += align_up(rows * C * sizeof(__half), 256);
    __half* hk = reinterpret_cast<__half*>(sp);
    sp += align_up(rows * C * sizeof(__half), 256);
    float* xn = nullptr;
    if (do_normalize) { xn = reinterpret_cast<float*>(sp); sp += align_up(rows * C * sizeof(float), 256); }
    int32_t* surv = reinterpret_cast<int32_t*>(sp);
    sp += align_up(rows * plan->kl * sizeof(int32_t), 256);
    int32_t* cnt = reinterpret_cast<int32_t*>(sp);
    sp += align_up(rows * sizeof(int32_t), 256);
    int32_t* ovf_list = reinterpret_cast<int32_t*>(sp);
    if (rows >= (1ull << 31)) return CRW_ERR_UNSUPPORTED;

    CRW_CUDA_RET(cudaMemsetAsync(stats, 0, head, st));
    if (rows > 0) {
        const int rows_rg = T * N;
        if (R > 65535) return CRW_ERR_UNSUPPORTED;
        // centre from a sample of the rows, then one pass over the features with a bounded grid (~8 CTAs per SM over all radargrams)
        lp_mu_x_kernel<<<dim3(R <= 8 ? 128u : 16u, (unsigned)R), 256, 0, st>>>(feats, rows_rg, do_normalize, musum);
        CRW_LAUNCH_RET();
        const unsigned per_rg = (unsigned)max(1, min((rows_rg + 15) / 16, (8 * max(sms, 1) + R - 1) / R));
        lp_prep_x_kernel<<<dim3(per_rg, (unsigned)R), 256, 0, st>>>(feats, rows_rg, do_normalize, xn, hq, hk, musum, stats);
        CRW_LAUNCH_RET();
    }
    XParams& p = plan->p;
    p.R = R; p.T = T; p.N = N; p.ctx = ctx; p.k = k;
    p.n_min = n_min < 1 ? 1 : n_min;
    p.rows_rg = T * N;
    const float rc_ = ceilf(radius);
    p.rb = (rc_ - 1.0f >= (float)N) ? N : (int)rc_ - 1;
    p.inv_temp = 1.0f / temp;
    p.magic_n = (unsigned)((1ull << 32) / (unsigned)N) + 1u;
    p.total_rows = (long long)rows;
    p.surv = surv; p.cnt = cnt; p.stats = stats;
    p.ovf_list = ovf_list; p.ovf_ctr = ovf_ctr;            // (lp_x_launch offsets both for the bulk launch)
    { const char* e = getenv("CRW_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
    // early items: the rows of frames 0 .. ctx+1 (the sequential gather waits for them) are cut into single query tiles, so that
    // this side job takes a fraction of a bulk item's time
    const int early_need = min((ctx + 2) * N, p.rows_rg);
    p.early_rpi = kXBM;
    p.early_items_rg = ceil_div(early_need, p.early_rpi);
    p.early_rows = min(p.early_items_rg * p.early_rpi, p.rows_rg);
    // bulk items: as many as fill whole rounds of the CTAs the bulk launch gets (all SMs but the ones the early items occupy
    // when those are few), at most kXG * 128 rows each; the stream of one item must fit the 12-bit column
    const int bulk_rows_rg = p.rows_rg - p.early_rows;
    int rpi = kXG * kXBM;
    if (bulk_rows_rg > 0) {
        const int max_rows = kXG * kXBM;
        const int early_ctas = R * p.early_items_rg;
        const int nb = (sms > 0) ? ((early_ctas <= sms / 8) ? sms - early_ctas : sms) : 0;
        long long items = ((long long)R * bulk_rows_rg + max_rows - 1) / max_rows;
        if (nb > 0) items = (items + nb - 1) / nb * nb;
        int items_rg = (int)((items + R - 1) / R);
        if (items_rg < 1) items_rg = 1;
        rpi = ceil_div(bulk_rows_rg, items_rg);
        if (rpi < 1) rpi = 1;
    }
    // stream length check: frame-0 tiles + (rows of the item + ctx + 1 frames) in 64-row tiles must stay below 2^12 / 64 tiles
    while (true) {
        const int frames = rpi / N + 2 + ctx;
        const int ktiles = ceil_div(N, kXBN) + ceil_div(frames * N, kXBN) + 1;
        if (ktiles <= (1 << kXColBits) / kXBN) break;
        if (rpi <= N) return CRW_ERR_UNSUPPORTED;
        rpi = rpi / 2;
    }
    p.rows_per_item = rpi;
    p.items_per_rg = p.early_items_rg + ceil_div(bulk_rows_rg, rpi);
    p.v_begin = p.v_end = 0;
    RParams& r = plan->r;
    r.xn = do_normalize ? xn : feats;
    r.surv = surv; r.cnt = cnt; r.W = W; r.I = I;
    r.ovf_list = ovf_list; r.ovf_ctr = ovf_ctr; r.resc_ctas = 0;
    r.R = R; r.T = T; r.N = N; r.ctx = ctx; r.rb = p.rb; r.k = k;
    r.n_min = p.n_min;
    r.rows_rg = p.rows_rg; r.row_begin = 0; r.row_end = 0;
    r.total_rows = p.total_rows; r.inv_temp = p.inv_temp; r.magic_n = p.magic_n; r.debug = p.debug;
    if (T < 2) return CRW_OK;
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(plan->maps);
    int rc = make_tmap_bf16_k64(&maps[0], hq, (uint64_t)rows, 128, kXBM);     // 2-byte elements: the bf16 map type moves fp16 as well
    if (rc == CRW_OK) rc = make_tmap_bf16_k64(&maps[1], hk, (uint64_t)rows, 128, kXBN);
    return rc;
}
int lp_x_total_slots(const void* plan) { const LpXPlan* pl = reinterpret_cast<const LpXPlan*>(plan); return pl->p.R * pl->p.items_per_rg; }
int lp_x_early_slots(const void* plan) { const LpXPlan* pl = reinterpret_cast<const LpXPlan*>(plan); return pl->p.R * pl->p.early_items_rg; }
int lp_x_early_rows(const void* plan) { return reinterpret_cast<const LpXPlan*>(plan)->p.early_rows; }

// filter over schedule slots [v_begin, v_end), then refine over the rows those slots cover (early slots: rows [0, early_rows) of
// every radargram; the rest: [early_rows, rows_rg)); v ranges must be exactly the early part, the rest, or everything
int lp_x_launch(const void* plan_storage, int v_begin, int v_end, int max_ctas, cudaStream_t st) {
    const LpXPlan& plan = *reinterpret_cast<const LpXPlan*>(plan_storage);
    if (v_end <= v_begin) return CRW_OK;
    XParams p = plan.p;
    p.v_begin = v_begin;
    p.v_end = v_end;
    // overflow list of this launch: the early launch fills [0, R * early_rows), anything else the rest (a launch over everything
    // starts at 0 as well: the two never run in one call)
    const bool bulk_part = v_begin > 0;
    const size_t list_off = bulk_part ? (size_t)plan.p.R * lp_x_early_rows(plan_storage) : 0;
    p.ovf_list = plan.p.ovf_list + list_off;
    p.ovf_ctr = plan.p.ovf_ctr + (bulk_part ? 2 : 0);
    int rc;
    if (plan.kt == 10) rc = launch_filter<10, 16>(plan, p, max_ctas, st);
    else if (plan.kt == 16) rc = launch_filter<16, 24>(plan, p, max_ctas, st);
    else rc = launch_filter<24, 32>(plan, p, max_ctas, st);
    if (rc != CRW_OK) return rc;
    RParams r = plan.r;
    r.ovf_list = p.ovf_list;
    r.ovf_ctr = p.ovf_ctr;
    const int early_slots = p.R * p.early_items_rg, early_rows = lp_x_early_rows(plan_storage);
    r.row_begin = (v_begin >= early_slots) ? early_rows : 0;
    r.row_end = (v_end <= early_slots) ? early_rows : p.rows_rg;
    if (plan.kl == 16) return launch_refine<16>(r, max_ctas, st);
    if (plan.kl == 24) return launch_refine<24>(r, max_ctas, st);
    return launch_refine<32>(r, max_ctas, st);
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
