// labelprop_tc.cu -- tensor-core label-propagation affinity + top-k for sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces the same reference code as labelprop_f32.cu (src/imported/maskedatt.py:151-175 with the radius
// mask of :232-245), for the "bf16 path" of BASELINE.json: operands are error-compensated bf16 pairs
// (x = hi + lo, dot ~ hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM), which keeps >= 99.9 % label
// agreement on near-collinear embeddings where plain bf16 does not.
//
//   lp_prep_bf16_kernel   F.normalize (pinned order, bit-identical to the fp32 path) + hi/lo split
//   lp_topk_tc_kernel     persistent, warp-specialised:
//       warp 8      TMA producer: a ring of 32 KB stages, SWIZZLE_128B -- 64-row key tiles and (TS form, the default) the two
//                   planes of the next query tile (128 consecutive node rows)
//       warp 9      tcgen05 issuer: 24 MMAs (3 passes x 8 k-steps) per key tile into one of 4 TMEM accumulators; query planes
//                   are copied from their ring slot into a TMEM query buffer with tcgen05.cp, in ring order
//       warps 0-7   epilogue: tcgen05.ld a row per thread (thread = query node), frame-window + radius-band
//                   predicate, running top-k in registers (sorted inserts on the FMA pipe, toplist_insert.inc); the two lists of
//                   a query (key tiles are dealt to two warp groups) are merged through smem, then softmax and the W / I stores
//   Work items (Sched): whole query tiles, and the tiles of the last, partial round cut in two by key range (lists handed over
//   through global memory).  lp_topk_pair_kernel: the same on 2-CTA clusters with cta_group::2 MMAs (opt-in).
// A query tile is 128 consecutive rows of the [T*N, C] feature matrix, so tiles are dense even though
// frames (N = 47..49 rows) straddle them; each thread derives its own frame / window from its row index.
#include <cstdlib>
#include "tc_common.cuh"

namespace crw {

// ------------------------------------------------------------------------------------------
// prep: normalise (optional) + split into bf16 hi / lo.  One warp per row, C == 128.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lp_prep_bf16_kernel(const float* __restrict__ x, int64_t rows, int do_normalize,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           int* __restrict__ zero, int n_zero) {
    if (blockIdx.x == 0)          // arrival counters of the top-k kernel's split tiles (same stream, runs before it)
        for (int i = threadIdx.x; i < n_zero; i += 256) zero[i] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* xr = x + row * 128;
    float v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = xr[lane + 32 * m];
    if (do_normalize) {
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
        const __nv_bfloat16 h = __float2bfloat16_rn(v[m]);
        hi[row * 128 + lane + 32 * m] = h;
        lo[row * 128 + lane + 32 * m] = __float2bfloat16_rn(v[m] - __bfloat162float(h));
    }
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
constexpr int kBM = 128;            // query rows per tile (TMEM lanes)
constexpr int kBN = 64;             // key rows per stage (TMEM columns per accumulator buffer)
constexpr int kStages = 3;          // key smem stages
constexpr int kAcc = 8;             // TMEM accumulator buffers (8 x 64 columns = all 512)
constexpr int kQBytes = 4 * kBM * 128;   // [hi,lo][kblock 0,1][128 rows][128 B] = 64 KB
constexpr int kKBytes = 4 * kBN * 128;   // [hi,lo][kblock 0,1][ 64 rows][128 B] = 32 KB
constexpr int kParkBytes = 64 * 1024;    // epilogue scratch, split evenly between the epilogue warps

struct TcParams {
    float* W;        // [R, T, k, N]  (row n of radargram r at ((r*T + n)*k + j)*N + q)
    int32_t* I;
    int R, T, N, ctx, rb, k;
    float inv_temp;
    int tiles_per_rg, total_tiles;
    int early_per_rg;   // leading tiles of every radargram that hold the query rows of frames 0..ctx+1 (scheduled first)
    int v_begin, v_end; // range of schedule slots this launch walks (slot -> tile: early tiles of all radargrams first)
    unsigned magic_n;   // floor(2^32 / N) + 1: x / N == __umulhi(x, magic_n) for x * N < 2^32
    int debug;          // profiling aid (env CRW_TC_DEBUG): 1 = skip insertions, 2 = also skip filter/park; results invalid
    const __nv_bfloat16* hi;   // [R*T*N, 128] operand planes
    const __nv_bfloat16* lo;
    long long total_rows;
    // tail split (see Sched): partial lists [tail tile][half][value s | id s][128 rows] and arrival counters [tail tile][quadrant]
    float* pbuf;
    int* pcnt;
    int split_ok;
};
constexpr int kMaxSplitTiles = 74;       // at most half the SMs of a B200 own a tile in the last, partial round

// host-side plan of one tensor-path call (opaque to labelprop_f32.cu: it only sees the size)
struct LpTcPlan {
    alignas(64) unsigned char maps[4 * sizeof(CUtensorMap)];
    TcParams p;
    int pair;
    int ts;
};

struct TileInfo {
    int rg, r0;          // radargram, first (radargram-relative) query row
    int n_lo, n_hi;      // first / last valid query frame in the tile (n_lo > n_hi: nothing to do)
    int f_lo;            // first key frame of the contiguous key range [f_lo*N, n_hi*N)
    int has_f0;          // number of separate frame-0 tiles preceding the contiguous range
    int n_ktiles;        // key tiles including the frame-0 tiles
};

// schedule slot -> (radargram, tile within it): the early tiles of all radargrams occupy the first R * early_per_rg slots so
// that the sequential part of the label gather (frames 1..ctx+1) can start while the rest of the top-k is still running
__device__ __forceinline__ void slot_to_tile(const TcParams& p, int v, int& rg, int& tt) {
    const int E = p.early_per_rg, early_total = p.R * E;
    if (v < early_total) { rg = v / E; tt = v - rg * E; return; }
    const int w = v - early_total, rest = p.tiles_per_rg - E;
    rg = w / rest;
    tt = E + (w - rg * rest);
}
__device__ __forceinline__ TileInfo tile_info(const TcParams& p, int tile) {
    TileInfo t;
    int tt;
    slot_to_tile(p, tile, t.rg, tt);
    t.r0 = tt * kBM;
    t.n_lo = max(1, t.r0 / p.N);
    t.n_hi = min(p.T - 1, (t.r0 + kBM - 1) / p.N);
    t.f_lo = max(0, t.n_lo - p.ctx);
    t.has_f0 = (t.f_lo > 0) ? ceil_div(p.N, kBN) : 0;
    t.n_ktiles = (t.n_lo > t.n_hi) ? 0 : t.has_f0 + ceil_div((t.n_hi - t.f_lo) * p.N, kBN);
    return t;
}
// first key row (radargram-relative) and number of valid key rows of key tile kt
__device__ __forceinline__ void ktile_rows(const TcParams& p, const TileInfo& t, int kt, int& row0, int& nrows) {
    if (kt < t.has_f0) { row0 = kt * kBN; nrows = min(kBN, p.N - row0); return; }
    const int c = kt - t.has_f0;
    row0 = t.f_lo * p.N + c * kBN;
    nrows = min(kBN, t.n_hi * p.N - row0);
}

// Work items of one launch.  A CTA walks items blockIdx.x, blockIdx.x + G, ...  Items are whole query tiles, except that the tiles of
// the last, partial round (tail = tiles mod G, when 2 tail <= G) are cut in two by KEY range: 2 tail items on 2 tail different CTAs,
// each producing a partial top-k list per query; the half that arrives second (a counter per tile and lane quadrant) merges the
// two lists and finishes the query.  470 tiles on 139 CTAs are 3 full rounds + 53 tiles = 106 half items instead of a 4th round.
struct Sched {
    int G, full, tail, n_items;
    bool split;
    __host__ __device__ __forceinline__ Sched(int n_tiles, int grid, bool split_ok) {
        G = grid;
        full = (n_tiles / G) * G;
        tail = n_tiles - full;
        split = split_ok && full > 0 && tail > 0 && 2 * tail <= G && tail <= kMaxSplitTiles;
        n_items = full + (split ? 2 * tail : tail);
    }
    __device__ __forceinline__ Sched(const TcParams& p, int grid) : Sched(p.v_end - p.v_begin, grid, p.split_ok != 0) {}
    // item -> tile (relative to the launch's first slot), half (-1: whole tile), index among the split tiles
    __host__ __device__ __forceinline__ void map(int item, int& tile_rel, int& half, int& tidx) const {
        tile_rel = item; half = -1; tidx = 0;
        if (split) {
            if (item < 2 * tail) {
                half = (item >= tail) ? 1 : 0;
                tidx = item - half * tail;
                tile_rel = full + tidx;
            } else {
                tile_rel = item - 2 * tail;
            }
        }
    }
    // item -> tile info, key-tile range [kt_lo, kt_hi), half (-1: whole tile) and index of the tile among the split ones.
    // The halves are the FIRST items (one per CTA, 2 tail CTAs): the hand-over at the end of a half (global stores, counter,
    // maybe the merge) then overlaps the start of the CTA's next item instead of sitting at the very end of the launch.
    __device__ __forceinline__ TileInfo decode(const TcParams& p, int item, int& kt_lo, int& kt_hi, int& half, int& tidx) const {
        int tile;
        map(item, tile, half, tidx);
        tile += p.v_begin;
        const TileInfo t = tile_info(p, tile);
        kt_lo = 0; kt_hi = t.n_ktiles;
        if (half == 0) kt_hi = t.n_ktiles / 2;
        if (half == 1) kt_lo = t.n_ktiles / 2;
        return t;
    }
    __device__ __forceinline__ bool has_work(const TcParams& p, int item) const {
        int a, b, h, x;
        return decode(p, item, a, b, h, x).n_ktiles != 0;
    }
    __device__ __forceinline__ int next(const TcParams& p, int item) const {      // next item of this CTA with work, or -1
        for (item += G; item < n_items; item += G)
            if (has_work(p, item)) return item;
        return -1;
    }
    __device__ __forceinline__ int first(const TcParams& p, int item) const {
        if (item < n_items && has_work(p, item)) return item;
        return next(p, item);
    }
};

// profiling aid (CRW_TC_DEBUG bit 3): cycles each epilogue warp spends per phase, summed over the launch
//   [0] waiting for an accumulator  [1] tcgen05.ld + park + threshold mask  [2] validity mask  [3] insertion loop
//   [4] merge + finish + stores     [5] insertion-loop iterations (count, not cycles)
//   inside [4]: [6] wait for the partner list (first barrier)  [7] list hand-over + merge  [8] finish (softmax, stores)
//               [9] last barrier (scratch free)
__device__ unsigned long long g_lp_prof[160 * 8 * 10];

// Sorted insert of (x, xid) into a descending list, one inline-PTX block per list length (generated, see
// tools/gen_toplist_insert.py): predicated FFMA moves on the FMA pipe instead of SEL / FSEL on the ALU pipe.
template <int KT>
__device__ __forceinline__ void toplist_insert_fma(float (&v)[KT], float (&idf)[KT], float x, float xid);
template <int KT>
__device__ __forceinline__ void toplist_insert_tie_fma(float (&v)[KT], float (&idf)[KT], float x, float xid);
#include "toplist_insert.inc"

constexpr float kIdNone = 16777216.0f;      // ids (key rows) travel as floats: exact below 2^24 (checked by lp_tc_prepare)
constexpr float kNoValue = -3.402823466e+38f;   // empty list slot: finite (the insert multiplies slots by 0), below every dot product

template <int KT>
struct TopList {
    float v[KT];
    float idf[KT];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < KT; ++s) { v[s] = kNoValue; idf[s] = kIdNone; }
    }
    // insert keeping (value desc); a later candidate never displaces an equal value
    __device__ __forceinline__ void insert(float x, float xid) { toplist_insert_fma<KT>(v, idf, x, xid); }
    // full comparator (value desc, id asc) for merging lists whose ids interleave
    __device__ __forceinline__ void insert_tie(float x, float xid) { toplist_insert_tie_fma<KT>(v, idf, x, xid); }
};

// Branch-free pop of the lowest set bit of the candidate mask (c1:c0) and fetch of that column from the lane's park
// slots.  Returns false (x = -inf, a no-op for insert()) when the mask is empty; the fetch then reads slot 31, in range.
// The id comes out as a float: (2^23 + i) + (row0 - 2^23), both terms and the sum exact.
__device__ __forceinline__ bool pop_candidate(uint32_t& c0, uint32_t& c1, uint32_t park, float row0_bias, float& x, float& xid) {
    const bool any = (c0 | c1) != 0u, in_lo = c0 != 0u;
    const uint32_t w = in_lo ? c0 : c1, nw = w & (w - 1u);
    const int i = (__ffs(w) - 1 + (in_lo ? 0 : 32)) & 63;
    c0 = in_lo ? nw : c0;
    c1 = in_lo ? c1 : nw;
    const float v = tc::lds_f32(park + i * 128);
    x = any ? v : -INFINITY;
    xid = __int_as_float(0x4B000000 | i) + row0_bias;
    return any;
}
__device__ __forceinline__ float row_bias(int row0) { return (float)row0 - 8388608.0f; }

// acc += 2^(i & 15) when v > thr: the "beats the threshold" bitmask of 32 parked values is collected as two exact float sums of
// distinct powers of two (16 bits each) with predicated FADDs on the FMA pipe -- the SEL + IADD3 form of `mask |= p << i`
// sits on the ALU pipe, which the selection epilogue saturates.
template <int I>
__device__ __forceinline__ void mask_acc(float& acc, float v, float thr) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.f32 %0, %0, %3;\n\t}" : "+f"(acc) : "f"(v), "f"(thr), "f"((float)(1u << (I & 15))));
}
__device__ __forceinline__ uint32_t mask_from_acc(float lo16, float hi16) {
    return (uint32_t)__float2int_rz(lo16) | ((uint32_t)__float2int_rz(hi16) << 16);
}
template <int I>
struct ParkLoop {
    static __device__ __forceinline__ void run(const float (&v)[32], uint32_t park_ch, float thr, float& a0, float& a1) {
        tc::sts_f32(park_ch + I * 128, v[I]);
        if (I < 16) mask_acc<I>(a0, v[I], thr); else mask_acc<I>(a1, v[I], thr);
        ParkLoop<I + 1>::run(v, park_ch, thr, a0, a1);
    }
};
template <>
struct ParkLoop<32> {
    static __device__ __forceinline__ void run(const float (&)[32], uint32_t, float, float&, float&) {}
};

// Validity mask of one thread over the (<= 64) key rows [row0, row0 + nrows) of a key tile: bit c is set when key row row0 + c
// lies in an allowed key frame (kf < n and kf == 0 or kf >= win_lo) and inside the radius band |node - q| <= rb.
// The band of the thread's query node is one run of bits (mq, first node lo_q); per key frame it is shifted to where that
// frame starts in the tile -- bits that leave the 64-bit word are exactly the nodes outside the tile.  Needs 2 rb + 1 <= 64.
__device__ __forceinline__ void band_mask(int row0, int nrows, int N, unsigned magic_n, int n, int win_lo, int lo_q,
                                          unsigned long long mq, uint32_t (&vm)[2]) {
    if (nrows <= 0) return;
    const int kf0 = (int)__umulhi((unsigned)row0, magic_n);
    unsigned long long m = 0ull;
    int kf = kf0;
    for (int o = kf0 * N - row0; o < nrows; o += N, ++kf) {          // warp-uniform trip count; o = tile column of node 0
        const int sh = o + lo_q;
        const unsigned long long seg = (sh >= 0) ? ((sh < 64) ? (mq << sh) : 0ull) : ((sh > -64) ? (mq >> (-sh)) : 0ull);
        m |= ((kf < n) && (kf == 0 || kf >= win_lo)) ? seg : 0ull;
    }
    if (nrows < 64) m &= (1ull << nrows) - 1ull;
    vm[0] = (uint32_t)m;
    vm[1] = (uint32_t)(m >> 32);
}

// a value strictly below x (within a few ulp); -inf stays -inf.  (nextafterf() is a ~20-instruction sequence.)
__device__ __forceinline__ float strictly_below(float x) { return __fmaf_rn(-fabsf(x), 2.384185791015625e-07f, x) - 1.17549435e-38f; }

template <int NEPI>
__device__ __forceinline__ float shared_threshold(const float* thr_pub, const float* mid_pub, int warp, int lane, float own,
                                                  float own_mid, bool use_mid = true) {
    float t = own;
#pragma unroll
    for (int pp = 1; pp < NEPI / 4; ++pp) {
        const float o = *reinterpret_cast<const volatile float*>(&thr_pub[((warp + 4 * pp) % NEPI) * 32 + lane]);
        t = fmaxf(t, strictly_below(o));
    }
    if (NEPI / 4 == 2 && use_mid) {
        const float m = *reinterpret_cast<const volatile float*>(&mid_pub[((warp + 4) % NEPI) * 32 + lane]);
        t = fmaxf(t, strictly_below(fminf(own_mid, m)));
    }
    return t;
}

// key rows -> candidate ids, fill with out-of-band candidates when fewer than k are in band, softmax, W / I stores
template <int KT>
__device__ __forceinline__ void lp_finish_query(const TcParams& p, TopList<KT>& top, int rg, int n, int q, int win_lo) {
    const int N = p.N, ctx = p.ctx, k = p.k, rb = p.rb;
    // key row -> candidate id (slot in the trimmed key set * N + node)
    int id[KT];
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        const int kr = (int)top.idf[s];
        const int kf = (int)__umulhi((unsigned)kr, p.magic_n), j = kr - kf * N;
        const int slot = (n > ctx + 1 && kf != 0) ? kf - win_lo + 1 : kf;
        id[s] = slot * N + j;
    }
    // fewer than k in-band candidates: out-of-band ones share one logit; ascending id (pinned tie rule)
    const int F = n_key_frames(n, ctx);
    int live = 0;
#pragma unroll
    for (int s = 0; s < KT; ++s) live += (s < k && top.v[s] > kNoValue) ? 1 : 0;
    const float masked = kMaskBias;   // raw-dot domain stand-in: exp() of it is exactly 0
    if (live < k) {
        int need = k - live, fill = live;
        for (int f = 0; f < F && need > 0; ++f)
            for (int jj = 0; jj < N && need > 0; ++jj) {
                const int dj = jj - q;
                if (dj <= rb && -dj <= rb) continue;
#pragma unroll
                for (int s = 0; s < KT; ++s)
                    if (s == fill) { top.v[s] = masked; id[s] = f * N + jj; }
                ++fill; --need;
            }
    }
    const float l0 = top.v[0] * p.inv_temp;
    float e[KT], sum = 0.0f;
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        e[s] = (s < k) ? pinned_expf(top.v[s] * p.inv_temp - l0) : 0.0f;
        sum += e[s];
    }
    const float inv = 1.0f / sum;
    const size_t base = ((size_t)(rg * p.T + n) * k) * N + q;
#pragma unroll
    for (int s = 0; s < KT; ++s)
        if (s < k) {
            p.W[base + (size_t)s * N] = e[s] * inv;
            p.I[base + (size_t)s * N] = id[s];
        }
}

// TS = true: the query tile is the A operand IN TENSOR MEMORY (tcgen05.mma "TS" form).  It travels through the key-stage ring:
// the producer loads each 32 KB plane (hi, lo) of the NEXT query tile into a free stage right after the first key tile of the
// current one, and the MMA warp, reaching those slots in the same order, copies them into one of two 128-column TMEM buffers
// with tcgen05.cp (ordered with the MMAs around it) and hands the stage back.  An MMA then reads only its 2 KB of B from shared
// memory instead of 6 KB, and the 64 KB of the shared-memory query tile hold two more key stages.  TMEM: [0,256) two query
// buffers (hi 64 | lo 64 columns each), [256,512) four 64-column accumulators.
template <int KT, int NEPI, bool TS>
__global__ void __launch_bounds__((NEPI + 2) * 32, 1)
lp_topk_tc_kernel(const __grid_constant__ CUtensorMap qmap_hi, const __grid_constant__ CUtensorMap qmap_lo,
                  const __grid_constant__ CUtensorMap kmap_hi, const __grid_constant__ CUtensorMap kmap_lo, TcParams p) {
    constexpr int kProducerWarp = NEPI, kMmaWarp = NEPI + 1;
    constexpr int kParts = NEPI / 4;          // top-k lists per query (merged at the end of a tile)
    constexpr int kParkWarp = kParkBytes / NEPI;
    constexpr int kNStages = TS ? (kQBytes + kStages * kKBytes) / kKBytes : kStages;   // same carve-up size: 5 stages
    constexpr int kNAcc = TS ? 4 : kAcc;
    constexpr uint32_t kAccCol0 = TS ? 256u : 0u;
    // 8 epilogue warps: two parts, whole 64-column key tiles dealt round-robin.  16 warps: four parts = two tile groups x two
    // 32-column halves (the park buffer of a warp then only has to hold 32 columns per lane).
    constexpr int kColSplit = (NEPI > 8) ? 2 : 1;
    constexpr int kTileGroups = kParts / kColSplit;
    constexpr int kCols = kBN / kColSplit, kCh = kCols / 32;
    static_assert(kParkWarp >= kCols * 32 * 4, "park buffer must hold this warp's columns of one key tile per lane");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // 64 KB (SS form only)
    uint8_t* sK = TS ? smem : smem + kQBytes;            // kNStages x 32 KB
    uint8_t* park_base = sK + kNStages * kKBytes;        // 64 KB
    {
        uint32_t dyn_bytes;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        if ((size_t)(park_base + kParkBytes - smem_raw) > dyn_bytes) __trap();    // the launch did not provide the carve-up
    }
    __shared__ uint64_t q_full[2], q_empty, k_full[kNStages], k_empty[kNStages], acc_full[kNAcc], acc_empty[kNAcc];
    __shared__ uint32_t tmem_base_s;
    __shared__ float thr_pub[NEPI * 32];     // every list's current k-th best, read by the other part(s) of the same query
    __shared__ float mid_pub[NEPI * 32];     // ... and its current ceil(KT/2)-th best (see shared_threshold)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;

    if (warp == kMmaWarp) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&q_full[0], 1);
        tc::mbar_init(&q_full[1], 1);
        tc::mbar_init(&q_empty, 1);
        for (int s = 0; s < kNStages; ++s) { tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); }
        for (int a = 0; a < kNAcc; ++a) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 4 * kColSplit); }
        tc::fence_barrier_init();
    }
    if (tid < NEPI * 32) { thr_pub[tid] = -INFINITY; mid_pub[tid] = -INFINITY; }
    if (warp == kProducerWarp && lane == 0) {
        tc::prefetch_tmap(&qmap_hi); tc::prefetch_tmap(&qmap_lo); tc::prefetch_tmap(&kmap_hi); tc::prefetch_tmap(&kmap_lo);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kProducerWarp) {
        // ================= TMA producer (whole warp runs the loop; one elected lane issues) =================
        const bool leader = tc::elect_one();
        uint32_t scnt = 0, tcnt = 0;          // ring slots handed out so far (key tiles and, TS, query planes); tiles done
        // TS: the two planes of query tile tl, one ring slot each ([kblock 0,1][128 rows][128 B])
        const Sched sched(p, gridDim.x);
        auto push_query = [&](int it) {
            int a_, b_, h_, x_;
            const TileInfo tq = sched.decode(p, it, a_, b_, h_, x_);
            const int grow_q = tq.rg * p.T * N + tq.r0;
            for (int plane = 0; plane < 2; ++plane, ++scnt) {
                const int s = scnt % kNStages;
                tc::mbar_wait_backoff(&k_empty[s], ((scnt / kNStages) & 1) ^ 1);
                if (leader) {
                    tc::mbar_arrive_expect_tx(&k_full[s], kKBytes);
                    uint8_t* dst = sK + s * kKBytes;
                    for (int kb = 0; kb < 2; ++kb)
                        tc::tma_load_2d(dst + kb * (kBM * 128), plane ? &qmap_lo : &qmap_hi, kb * 64, grow_q, &k_full[s]);
                }
            }
        };
        if (TS) {
            const int first = sched.first(p, blockIdx.x);
            if (first >= 0) push_query(first);
        }
        for (int item = blockIdx.x; item < sched.n_items; item += gridDim.x) {
            int kt_lo, kt_hi, half, tidx;
            const TileInfo t = sched.decode(p, item, kt_lo, kt_hi, half, tidx);
            if (t.n_ktiles == 0) continue;
            const int grow = t.rg * p.T * N;   // first global row of this radargram
            const int nx = TS ? sched.next(p, item) : -1;
            if (!TS) {
                tc::mbar_wait_backoff(&q_empty, (tcnt & 1) ^ 1);
                if (leader) {
                    tc::mbar_arrive_expect_tx(&q_full[0], kQBytes);
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub)
                        tc::tma_load_2d(sQ + sub * (kBM * 128), (sub & 2) ? &qmap_lo : &qmap_hi, (sub & 1) * 64, grow + t.r0, &q_full[0]);
                }
            }
            if (TS && kt_lo == kt_hi && nx >= 0) push_query(nx);      // (a half without key tiles still passes the next query on)
            for (int kt = kt_lo; kt < kt_hi; ++kt) {
                const int s = scnt % kNStages;
                int row0, nrows;
                ktile_rows(p, t, kt, row0, nrows);
                tc::mbar_wait_backoff(&k_empty[s], ((scnt / kNStages) & 1) ^ 1);
                if (leader) {
                    tc::mbar_arrive_expect_tx(&k_full[s], kKBytes);
                    uint8_t* dst = sK + s * kKBytes;
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub)
                        tc::tma_load_2d(dst + sub * (kBN * 128), (sub & 2) ? &kmap_lo : &kmap_hi, (sub & 1) * 64, grow + row0, &k_full[s]);
                }
                ++scnt;
                if (TS && kt == kt_lo && nx >= 0) push_query(nx);
            }
            ++tcnt;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
        const bool leader = tc::elect_one();
        const uint64_t qdesc = tc::umma_smem_desc_k128(tc::smem_u32(sQ));     // descriptor of sub-tile 0, k-step 0
        const uint64_t kdesc0 = tc::umma_smem_desc_k128(tc::smem_u32(sK));
        uint32_t kcnt = 0, scnt = 0, tcnt = 0;      // key tiles (-> accumulator), ring slots, tiles
        // TS: copy the two query planes waiting in the ring into TMEM query buffer buf (8 x 128x256b per plane)
        auto copy_query = [&](uint32_t buf) {
            for (int plane = 0; plane < 2; ++plane, ++scnt) {
                const int s = scnt % kNStages;
                tc::mbar_wait(&k_full[s], (scnt / kNStages) & 1);
                tc::tc_fence_after();
                if (leader) {
                    const uint64_t sd = kdesc0 + (uint64_t)((s * kKBytes) >> 4);
                    for (int kb = 0; kb < 2; ++kb)
                        for (int ks = 0; ks < 4; ++ks)
                            tc::tmem_cp_128x256b(tmem_base + buf * 128u + (uint32_t)(plane * 64 + kb * 32 + ks * 8),
                                                 sd + (uint64_t)((kb * (kBM * 128) + ks * 32) >> 4));
                    tc::umma_commit(&k_empty[s]);
                }
                __syncwarp();
            }
        };
        const Sched sched(p, gridDim.x);
        if (TS && sched.first(p, blockIdx.x) >= 0) copy_query(0u);
        for (int item = blockIdx.x; item < sched.n_items; item += gridDim.x) {
            int kt_lo, kt_hi, half, tidx;
            const TileInfo t = sched.decode(p, item, kt_lo, kt_hi, half, tidx);
            if (t.n_ktiles == 0) continue;
            const int nx = TS ? sched.next(p, item) : -1;
            if (!TS) tc::mbar_wait_backoff(&q_full[0], tcnt & 1);
            const uint32_t qtmem = tmem_base + (uint32_t)((tcnt & 1) * 128);      // TS: this tile's query buffer
            if (TS && kt_lo == kt_hi && nx >= 0) copy_query((tcnt + 1) & 1u);
            for (int kt = kt_lo; kt < kt_hi; ++kt, ++kcnt) {
                const int s = scnt % kNStages, a = kcnt % kNAcc;
                int row0, nrows;
                ktile_rows(p, t, kt, row0, nrows);
                const int ncols = min(kBN, (nrows + 15) & ~15);
                tc::mbar_wait(&k_full[s], (scnt / kNStages) & 1);           // latency critical: no backoff
                tc::mbar_wait_backoff(&acc_empty[a], ((kcnt / kNAcc) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t idesc = tc::umma_idesc_bf16(kBM, ncols);
                const uint64_t kdesc = kdesc0 + (uint64_t)((s * kKBytes) >> 4);   // start-address field counts 16-byte units
                const uint32_t d = tmem_base + kAccCol0 + (uint32_t)(a * kBN);
                if (leader) {
                    // pass 0: q_hi.k_hi   pass 1: q_hi.k_lo   pass 2: q_lo.k_hi ; sub-tile = part*2 + kblock, k-step = 32 bytes
                    if (!(p.debug & 4))
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        if ((p.debug & 128) && pass) break;      // timing aid: hi.hi pass only (results invalid)
                        const int qpart = (pass == 2) ? 2 : 0, kpart = (pass == 1) ? 2 : 0;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t bd = kdesc + (uint64_t)((((kpart + kb) * (kBN * 128)) + ks * 32) >> 4);
                                if (TS) {
                                    // 32-bit column c of a query buffer = K elements 2c, 2c+1: hi at [0,64), lo at [64,128)
                                    tc::umma_bf16_ts(d, qtmem + (uint32_t)((qpart ? 64 : 0) + kb * 32 + ks * 8), bd, idesc, (pass | kb | ks) ? 1u : 0u);
                                } else {
                                    const uint64_t ad = qdesc + (uint64_t)((((qpart + kb) * (kBM * 128)) + ks * 32) >> 4);
                                    tc::umma_bf16_ss(d, ad, bd, idesc, (pass | kb | ks) ? 1u : 0u);
                                }
                            }
                    }
                    // (timing aid, debug bit 8 of the second byte = 256, results INVALID: hand the stage back at once, so the key
                    // stream never waits for MMAs -- tells a dependency between the two apart from a shared resource)
                    if (p.debug & 256) tc::mbar_arrive(&k_empty[s]);
                    else tc::umma_commit(&k_empty[s]);     // smem stage may be refilled once these MMAs retire
                    tc::umma_commit(&acc_full[a]);    // accumulator ready for the epilogue
                }
                __syncwarp();
                ++scnt;
                if (TS && kt == kt_lo && nx >= 0) copy_query((tcnt + 1) & 1u);   // the next item's query planes follow in the ring
            }
            if (!TS && leader) tc::umma_commit(&q_empty);    // query tile may be overwritten
            __syncwarp();
            ++tcnt;
        }
    } else {
        // ================= epilogue: NEPI warps; thread = query row (TMEM lane) =================
        // warp = 4*part + lane-group; key tiles (64 columns) are dealt round-robin to the parts, so every query has
        // kParts partial top-k lists that are merged at the end of the query tile.  Per key tile: (1) tcgen05.ld the
        // thread's 2 x 32 columns, (2) park them in a
        // thread-private smem column ([i][lane]: bank = lane, conflict-free) so candidates can be fetched by dynamic
        // index, (3) "beats the current k-th best" bitmask & validity bitmask (frame window x radius band),
        // (4) ONE rolled insertion loop over the set bits (keeps the hot loop inside the I-cache).
        const int g = warp & 3, part = warp >> 2;
        const int lrow = g * 32 + lane;
        const int rb = p.rb, ctx = p.ctx, k = p.k;
        const uint32_t park = tc::smem_u32(park_base + warp * kParkWarp) + lane * 4;
        uint32_t kcnt = 0;
        bool scratch_pending = false;     // part 1: part 0 may still be reading last tile's list out of this warp's park buffer
        const Sched sched(p, gridDim.x);
        for (int item = blockIdx.x; item < sched.n_items; item += gridDim.x) {
            int kt_lo, kt_hi, half, tidx;
            const TileInfo t = sched.decode(p, item, kt_lo, kt_hi, half, tidx);
            if (t.n_ktiles == 0) continue;
            const int row = t.r0 + lrow;
            const int n = row / N, q = row - n * N;
            const bool qvalid = (n >= 1) && (n < p.T);
            const int win_lo = (n > ctx + 1) ? n - ctx : 1;     // non-zero key frames allowed: [win_lo, n)
            const int lo_q = max(0, q - rb), w_q = min(N - 1, q + rb) - lo_q + 1;
            const unsigned long long mq = (w_q >= 64) ? ~0ull : ((1ull << w_q) - 1ull);   // the band of this query node
            const bool wide_band = 2 * rb + 1 > 64;             // (warp-uniform) band wider than a key tile: generic loop
            TopList<KT> top;                                    // ids hold the key ROW until the end of the tile
            top.init();
            for (int kt = kt_lo; kt < kt_hi; ++kt, ++kcnt) {
                // key tiles are dealt round-robin to the tile groups; the group that also merges and finishes the query (part 0)
                // takes the later residue, i.e. the smaller share when the count does not divide.  (Giving part 1 a few key
                // tiles more, to fill the time part 0 spends merging, measured slower: 114.0 / 114.6 / 115.3 / 116.8 us for 0-3.)
                if ((kt % kTileGroups) != kTileGroups - 1 - part / kColSplit) continue;
                const int a = kcnt % kNAcc;
                int row0, nrows;
                ktile_rows(p, t, kt, row0, nrows);
                const int col0 = (part % kColSplit) * kCols;                 // this warp's columns of the tile
                row0 += col0;
                nrows = min(kCols, nrows - col0);                            // <= 0: nothing in this half, only release the buffer
                const bool prof = (p.debug & 8) != 0;
                long long c_0 = prof ? clock64() : 0;
                tc::mbar_wait(&acc_full[a], (kcnt / kNAcc) & 1);
                tc::tc_fence_after();
                long long c_1 = prof ? clock64() : 0;
                const float thr = shared_threshold<NEPI>(thr_pub, mid_pub, warp, lane, top.v[KT - 1], top.v[(KT + 1) / 2 - 1], !(p.debug & 16));
                uint32_t pm[2] = {0u, 0u}, vm[2] = {0u, 0u};
                {
                    // both 32-column loads are in flight before the one wait; once the values sit in registers the TMEM
                    // buffer is handed back to the MMA warp (the insertion loop below reads the parked copy only)
                    const bool legacy = (p.debug & 32) != 0;      // A/B aid: wait per load, release after the insertions
                    float v[kCh][32];
#pragma unroll
                    for (int ch = 0; ch < kCh; ++ch)
                        if (ch * 32 < nrows && !(p.debug & 2)) {                    // warp-uniform
                            tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(g * 32) << 16) + kAccCol0 + (uint32_t)(a * kBN + col0 + ch * 32), v[ch]);
                            if (legacy) tc::tmem_ld_wait();
                        }
                    tc::tmem_ld_wait();
                    if (!legacy) {
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&acc_empty[a]);
                    }
                    if (scratch_pending) {          // (part 1 only, warp-uniform) once per query tile
                        asm volatile("bar.sync %0, %1;" ::"r"(5 + g), "n"(64) : "memory");
                        scratch_pending = false;
                    }
#pragma unroll
                    for (int ch = 0; ch < kCh; ++ch)
                        if (ch * 32 < nrows && !(p.debug & 2)) {
                            float a0 = 0.0f, a1 = 0.0f;
                            ParkLoop<0>::run(v[ch], park + ch * 32 * 128, thr, a0, a1);
                            pm[ch] = mask_from_acc(a0, a1);
                        }
                }
                long long c_2 = prof ? clock64() : 0;
                // validity mask of this thread over the tile's key rows (segments = key frames)
                if (!wide_band) {
                    band_mask(row0, nrows, N, p.magic_n, n, win_lo, lo_q, mq, vm);
                } else {
                    const int kf0 = (int)__umulhi((unsigned)row0, p.magic_n);
                    int c = 0, kf = kf0, j = row0 - kf0 * N;
                    while (c < nrows) {                           // warp-uniform trip count
                        const int seg = min(nrows - c, N - j);
                        if ((kf < n) && (kf == 0 || kf >= win_lo)) {
                            const int lo = max(j, q - rb), hi = min(j + seg - 1, q + rb);
                            if (lo <= hi) {
                                const int b0 = c + lo - j, nb = hi - lo + 1;
                                const unsigned long long m64 = ((nb >= 64) ? ~0ull : ((1ull << nb) - 1ull)) << b0;
                                vm[0] |= (uint32_t)m64;
                                vm[1] |= (uint32_t)(m64 >> 32);
                            }
                        }
                        c += seg; j = 0; ++kf;
                    }
                }
                uint32_t c0 = qvalid ? (pm[0] & vm[0]) : 0u, c1 = qvalid ? (pm[1] & vm[1]) : 0u;
                if (p.debug & 3) { c0 = 0; c1 = 0; }
                long long c_3 = prof ? clock64() : 0;
                int iters = 0;
                {
                    // software-pipelined: the next candidate is popped and fetched while the current one is inserted
                    const float rbias = row_bias(row0);
                    float x, xid;
                    bool have = pop_candidate(c0, c1, park, rbias, x, xid);
                    while (have) {                   // two candidates per trip: no register rotation between trips
                        float xn, xidn;
                        pop_candidate(c0, c1, park, rbias, xn, xidn);
                        top.insert(x, xid);
                        have = pop_candidate(c0, c1, park, rbias, x, xid);
                        top.insert(xn, xidn);        // x = -inf (mask empty): a no-op
                        if (prof) iters += 2;
                    }
                }
                thr_pub[warp * 32 + lane] = top.v[KT - 1];
                mid_pub[warp * 32 + lane] = top.v[(KT + 1) / 2 - 1];
                if (p.debug & 32) {
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[a]);
                }
                if (prof) {
                    const long long c_4 = clock64();
                    iters = __reduce_max_sync(0xffffffffu, iters);
                    if (lane == 0) {
                        unsigned long long* g = g_lp_prof + ((size_t)blockIdx.x * 8 + warp) * 10;
                        g[0] += c_1 - c_0; g[1] += c_2 - c_1; g[2] += c_3 - c_2; g[3] += c_4 - c_3; g[5] += iters;
                    }
                }
            }
            // ---- merge the kParts lists of every query (warps part>0 -> smem -> warp part 0) ----
            // Two parts: part 1 hands its list over and goes on to the next query tile at once; part 0 tells it through a second
            // named barrier (arrive / sync) when the list has been read, which part 1 only needs before it parks values again.
            constexpr bool kRunAhead = (kParts == 2);
            const bool profm = (p.debug & 8) != 0;
            const long long c_m = profm ? clock64() : 0;
            if (kRunAhead && scratch_pending) {       // a query tile without key tiles for this part: settle the hand-shake now
                asm volatile("bar.sync %0, %1;" ::"r"(5 + g), "n"(64) : "memory");
                scratch_pending = false;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");   // every epilogue warp is done with its park buffer
            const long long c_m1 = profm ? clock64() : 0;
            float* mv = reinterpret_cast<float*>(park_base + warp * kParkWarp);
            float* mi = mv + KT * 32;
            static_assert(KT * 32 * 8 <= kParkWarp, "merge scratch must fit the warp's park buffer");
            if (part != 0) {
#pragma unroll
                for (int s = 0; s < KT; ++s) { mv[s * 32 + lane] = top.v[s]; mi[s * 32 + lane] = top.idf[s]; }
            }
            // published bounds belong to this query tile: cleared before anyone can start the next one
            thr_pub[warp * 32 + lane] = -INFINITY;
            mid_pub[warp * 32 + lane] = -INFINITY;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");
            long long c_m2 = c_m1;
            if (part == 0) {
                for (int pp = 1; pp < kParts; ++pp) {
                    const float* pv = reinterpret_cast<const float*>(park_base + (pp * 4 + g) * kParkWarp);
                    const float* pi = pv + KT * 32;
                    // (an empty slot, kNoValue / kIdNone, is a no-op for the comparator; lists are sorted, so once no lane's
                    // entry reaches its current k-th best the rest cannot either)
#pragma unroll 2
                    for (int s = 0; s < KT; ++s) {
                        const float x = pv[s * 32 + lane];
                        if (!__any_sync(0xffffffffu, x >= top.v[KT - 1])) break;
                        top.insert_tie(x, pi[s * 32 + lane]);
                    }
                }
                if (kRunAhead) asm volatile("bar.arrive %0, %1;" ::"r"(5 + g), "n"(64) : "memory");   // part 1's buffer is free again
                c_m2 = profm ? clock64() : 0;
                bool finish = true;
                if (half >= 0) {
                    // half of a split tile: publish this half's list, count the arrival; the second half to arrive merges both
                    float* mine = p.pbuf + ((size_t)(tidx * 2 + half) * 2 * KT) * kBM + lrow;
#pragma unroll
                    for (int s = 0; s < KT; ++s) { mine[s * kBM] = top.v[s]; mine[(KT + s) * kBM] = top.idf[s]; }
                    // the lanes' stores happen before lane 0's release (ordered by the warp barrier before it); its acquire orders the
                    // other half's list before the loads below (ordered by the warp barrier after it)
                    __syncwarp();
                    int arrived = 0;
                    if (lane == 0)
                        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(arrived) : "l"(&p.pcnt[tidx * 4 + g]) : "memory");
                    arrived = __shfl_sync(0xffffffffu, arrived, 0);
                    __syncwarp();
                    finish = arrived == 1;
                    if (finish) {
                        const float* other = p.pbuf + ((size_t)(tidx * 2 + (1 - half)) * 2 * KT) * kBM + lrow;
#pragma unroll 2
                        for (int s = 0; s < KT; ++s) {
                            const float x = __ldcg(other + s * kBM);
                            if (!__any_sync(0xffffffffu, x >= top.v[KT - 1])) break;
                            top.insert_tie(x, __ldcg(other + (KT + s) * kBM));
                        }
                    }
                }
                if (finish && qvalid) lp_finish_query<KT>(p, top, t.rg, n, q, win_lo);
            } else {
                if (kRunAhead) scratch_pending = true;
                c_m2 = profm ? clock64() : 0;
            }
            const long long c_m3 = profm ? clock64() : 0;
            if (!kRunAhead) asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");   // scratch free for the next tile
            if (profm && lane == 0) {
                unsigned long long* gp = g_lp_prof + ((size_t)blockIdx.x * 8 + warp) * 10;
                const long long c_e = clock64();
                gp[4] += c_e - c_m; gp[6] += c_m1 - c_m; gp[7] += c_m2 - c_m1; gp[8] += c_m3 - c_m2; gp[9] += c_e - c_m3;
            }
        }
        if (scratch_pending) asm volatile("bar.sync %0, %1;" ::"r"(5 + g), "n"(64) : "memory");   // leave no barrier half-armed
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2 on 2-CTA clusters): one pair owns 256 consecutive query rows (128 per CTA) and
// walks the union key window in 128-row key tiles, 64 rows loaded by each CTA.  Per key row streamed from L2 the pair
// serves twice as many queries, and each SM reads only its own half of the B operand from shared memory, which is what
// bounded the single-CTA kernel (profiles/r01_lp_tc_anatomy.txt).  Roles per CTA as above; only the leader's warp 9
// issues MMAs, its commits are multicast to both CTAs' barriers, and the peer's epilogue warps release accumulator
// buffers on the leader's barriers.
// ------------------------------------------------------------------------------------------
constexpr int kPairM = 256;         // query rows per pair tile
constexpr int kPairN = 128;         // key rows per pair key tile (TMEM columns per accumulator buffer)
constexpr int kPairAcc = 4;         // 4 x 128 columns = all 512

__device__ __forceinline__ TileInfo pair_tile_info(const TcParams& p, int tile) {
    TileInfo t;
    int tt;
    slot_to_tile(p, tile, t.rg, tt);
    t.r0 = tt * kPairM;
    t.n_lo = max(1, t.r0 / p.N);
    t.n_hi = min(p.T - 1, (t.r0 + kPairM - 1) / p.N);
    t.f_lo = max(0, t.n_lo - p.ctx);
    t.has_f0 = (t.f_lo > 0) ? ceil_div(p.N, kPairN) : 0;
    t.n_ktiles = (t.n_lo > t.n_hi) ? 0 : t.has_f0 + ceil_div((t.n_hi - t.f_lo) * p.N, kPairN);
    return t;
}
__device__ __forceinline__ void pair_ktile_rows(const TcParams& p, const TileInfo& t, int kt, int& row0, int& nrows) {
    if (kt < t.has_f0) { row0 = kt * kPairN; nrows = min(kPairN, p.N - row0); return; }
    const int c = kt - t.has_f0;
    row0 = t.f_lo * p.N + c * kPairN;
    nrows = min(kPairN, t.n_hi * p.N - row0);
}

template <int KT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
lp_topk_pair_kernel(const __grid_constant__ CUtensorMap qmap_hi, const __grid_constant__ CUtensorMap qmap_lo,
                    const __grid_constant__ CUtensorMap kmap_hi, const __grid_constant__ CUtensorMap kmap_lo, TcParams p) {
    constexpr int NEPI = 8, kProducerWarp = NEPI, kMmaWarp = NEPI + 1;
    constexpr int kParkWarp = kParkBytes / NEPI;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // 64 KB: this CTA's 128 query rows
    uint8_t* sK = smem + kQBytes;                        // kStages x 32 KB: this CTA's 64 rows of each key tile
    uint8_t* park_base = sK + kStages * kKBytes;         // 64 KB
    {
        uint32_t dyn_bytes;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        if ((size_t)(park_base + kParkBytes - smem_raw) > dyn_bytes) __trap();    // the launch did not provide the carve-up
    }
    __shared__ uint64_t q_full, q_empty, k_full[kStages], k_empty[kStages], acc_full[kPairAcc], acc_empty[kPairAcc];
    __shared__ uint32_t tmem_base_s;
    __shared__ float thr_pub[NEPI * 32];
    __shared__ float mid_pub[NEPI * 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == kMmaWarp) tc::tmem_alloc_pair<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&q_full, 1);
        tc::mbar_init(&q_empty, 1);
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); }
        for (int a = 0; a < kPairAcc; ++a) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 2 * NEPI); }
        tc::fence_barrier_init();
    }
    if (tid < NEPI * 32) { thr_pub[tid] = -INFINITY; mid_pub[tid] = -INFINITY; }
    if (warp == kProducerWarp && lane == 0) {
        tc::prefetch_tmap(&qmap_hi); tc::prefetch_tmap(&qmap_lo); tc::prefetch_tmap(&kmap_hi); tc::prefetch_tmap(&kmap_lo);
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kProducerWarp) {
        // ================= TMA producer (both CTAs; the leader also arms the barriers with both CTAs' bytes) =================
        const bool leader_lane = tc::elect_one();
        const uint32_t q_full_l = tc::mapa_u32(tc::smem_u32(&q_full), 0);
        uint32_t kcnt = 0, tcnt = 0;
        for (int tile = p.v_begin + cluster_id; tile < p.v_end; tile += n_clusters) {
            const TileInfo t = pair_tile_info(p, tile);
            if (t.n_ktiles == 0) continue;
            const int grow = t.rg * p.T * N;
            tc::mbar_wait_backoff(&q_empty, (tcnt & 1) ^ 1);
            if (leader_lane) {
                if (rank == 0) tc::mbar_arrive_expect_tx(&q_full, 2 * kQBytes);
#pragma unroll
                for (int sub = 0; sub < 4; ++sub)
                    tc::tma_load_2d_pair(sQ + sub * (kBM * 128), (sub & 2) ? &qmap_lo : &qmap_hi, (sub & 1) * 64,
                                         grow + t.r0 + (int)rank * kBM, q_full_l);
            }
            for (int kt = 0; kt < t.n_ktiles; ++kt, ++kcnt) {
                const int s = kcnt % kStages;
                int row0, nrows;
                pair_ktile_rows(p, t, kt, row0, nrows);
                tc::mbar_wait_backoff(&k_empty[s], ((kcnt / kStages) & 1) ^ 1);
                if (leader_lane) {
                    const uint32_t k_full_l = tc::mapa_u32(tc::smem_u32(&k_full[s]), 0);
                    if (rank == 0) tc::mbar_arrive_expect_tx(&k_full[s], 2 * kKBytes);
                    uint8_t* dst = sK + s * kKBytes;
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub)
                        tc::tma_load_2d_pair(dst + sub * (kBN * 128), (sub & 2) ? &kmap_lo : &kmap_hi, (sub & 1) * 64,
                                             grow + row0 + (int)rank * kBN, k_full_l);
                }
            }
            ++tcnt;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer: leader CTA only =================
        if (rank == 0) {
            const bool leader_lane = tc::elect_one();
            const uint64_t qdesc = tc::umma_smem_desc_k128(tc::smem_u32(sQ));
            const uint64_t kdesc0 = tc::umma_smem_desc_k128(tc::smem_u32(sK));
            const uint32_t idesc = tc::umma_idesc_bf16(kPairM, kPairN);
            uint32_t kcnt = 0, tcnt = 0;
            for (int tile = p.v_begin + cluster_id; tile < p.v_end; tile += n_clusters) {
                const TileInfo t = pair_tile_info(p, tile);
                if (t.n_ktiles == 0) continue;
                tc::mbar_wait_backoff(&q_full, tcnt & 1);
                for (int kt = 0; kt < t.n_ktiles; ++kt, ++kcnt) {
                    const int s = kcnt % kStages, a = kcnt % kPairAcc;
                    tc::mbar_wait(&k_full[s], (kcnt / kStages) & 1);
                    tc::mbar_wait_backoff(&acc_empty[a], ((kcnt / kPairAcc) & 1) ^ 1);
                    tc::tc_fence_after();
                    const uint64_t kdesc = kdesc0 + (uint64_t)((s * kKBytes) >> 4);
                    const uint32_t d = tmem_base + (uint32_t)(a * kPairN);
                    if (leader_lane) {
                        if (!(p.debug & 4))
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const int qpart = (pass == 2) ? 2 : 0, kpart = (pass == 1) ? 2 : 0;
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint64_t ad = qdesc + (uint64_t)((((qpart + kb) * (kBM * 128)) + ks * 32) >> 4);
                                    const uint64_t bd = kdesc + (uint64_t)((((kpart + kb) * (kBN * 128)) + ks * 32) >> 4);
                                    tc::umma_bf16_ss_pair(d, ad, bd, idesc, (pass | kb | ks) ? 1u : 0u);
                                }
                        }
                        tc::umma_commit_pair(&k_empty[s], 0b11);
                        tc::umma_commit_pair(&acc_full[a], 0b11);
                    }
                    __syncwarp();
                }
                if (leader_lane) tc::umma_commit_pair(&q_empty, 0b11);
                __syncwarp();
                ++tcnt;
            }
        }
    } else {
        // ================= epilogue: 8 warps; thread = query row (TMEM lane); part p = columns [64p, 64p+64) of every key tile =====
        const int g = warp & 3, part = warp >> 2;
        const int lrow = (int)rank * kBM + g * 32 + lane;
        const int rb = p.rb, ctx = p.ctx, k = p.k;
        const uint32_t park = tc::smem_u32(park_base + warp * kParkWarp) + lane * 4;
        uint32_t kcnt = 0;
        for (int tile = p.v_begin + cluster_id; tile < p.v_end; tile += n_clusters) {
            const TileInfo t = pair_tile_info(p, tile);
            if (t.n_ktiles == 0) continue;
            const int row = t.r0 + lrow;
            const int n = row / N, q = row - n * N;
            const bool qvalid = (n >= 1) && (n < p.T);
            const int win_lo = (n > ctx + 1) ? n - ctx : 1;
            const int lo_q = max(0, q - rb), w_q = min(N - 1, q + rb) - lo_q + 1;
            const unsigned long long mq = (w_q >= 64) ? ~0ull : ((1ull << w_q) - 1ull);
            const bool wide_band = 2 * rb + 1 > 64;
            TopList<KT> top;
            top.init();
            for (int kt = 0; kt < t.n_ktiles; ++kt, ++kcnt) {
                const int a = kcnt % kPairAcc;
                int row0, nrows;
                pair_ktile_rows(p, t, kt, row0, nrows);
                row0 += part * kBN;                                   // this part's 64-column half
                nrows = min(kBN, nrows - part * kBN);
                tc::mbar_wait(&acc_full[a], (kcnt / kPairAcc) & 1);
                tc::tc_fence_after();
                if (nrows > 0) {
                    const float thr = shared_threshold<NEPI>(thr_pub, mid_pub, warp, lane, top.v[KT - 1], top.v[(KT + 1) / 2 - 1], !(p.debug & 16));
                    uint32_t pm[2] = {0u, 0u}, vm[2] = {0u, 0u};
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        if (ch * 32 < nrows && !(p.debug & 2)) {
                            float v[32];
                            tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(a * kPairN + part * kBN + ch * 32), v);
                            tc::tmem_ld_wait();
                            float a0 = 0.0f, a1 = 0.0f;
                            ParkLoop<0>::run(v, park + ch * 32 * 128, thr, a0, a1);
                            pm[ch] = mask_from_acc(a0, a1);
                        }
                    }
                    if (!wide_band) {
                        band_mask(row0, nrows, N, p.magic_n, n, win_lo, lo_q, mq, vm);
                    } else {
                        const int kf0 = (int)__umulhi((unsigned)row0, p.magic_n);
                        int c = 0, kf = kf0, j = row0 - kf0 * N;
                        while (c < nrows) {
                            const int seg = min(nrows - c, N - j);
                            if ((kf < n) && (kf == 0 || kf >= win_lo)) {
                                const int lo = max(j, q - rb), hi = min(j + seg - 1, q + rb);
                                if (lo <= hi) {
                                    const int b0 = c + lo - j, nb = hi - lo + 1;
                                    const unsigned long long m64 = ((nb >= 64) ? ~0ull : ((1ull << nb) - 1ull)) << b0;
                                    vm[0] |= (uint32_t)m64;
                                    vm[1] |= (uint32_t)(m64 >> 32);
                                }
                            }
                            c += seg; j = 0; ++kf;
                        }
                    }
                    uint32_t c0 = qvalid ? (pm[0] & vm[0]) : 0u, c1 = qvalid ? (pm[1] & vm[1]) : 0u;
                    if (p.debug & 3) { c0 = 0; c1 = 0; }
                    {
                        const float rbias = row_bias(row0);
                        float x, xid;
                        bool have = pop_candidate(c0, c1, park, rbias, x, xid);
                        while (have) {                   // two candidates per trip: no register rotation between trips
                            float xn, xidn;
                            pop_candidate(c0, c1, park, rbias, xn, xidn);
                            top.insert(x, xid);
                            have = pop_candidate(c0, c1, park, rbias, x, xid);
                            top.insert(xn, xidn);        // x = -inf (mask empty): a no-op
                        }
                    }
                }
                thr_pub[warp * 32 + lane] = top.v[KT - 1];
                mid_pub[warp * 32 + lane] = top.v[(KT + 1) / 2 - 1];
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(tc::smem_u32(&acc_empty[a]), 0));
            }
            // ---- merge the two column-half lists of every query (warps 4-7 -> smem -> warps 0-3) ----
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");
            float* mv = reinterpret_cast<float*>(park_base + warp * kParkWarp);
            float* mi = mv + KT * 32;
            if (part != 0) {
#pragma unroll
                for (int s = 0; s < KT; ++s) { mv[s * 32 + lane] = top.v[s]; mi[s * 32 + lane] = top.idf[s]; }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");
            if (part == 0) {
                const float* pv = reinterpret_cast<const float*>(park_base + (4 + g) * kParkWarp);
                const float* pi = pv + KT * 32;
#pragma unroll 2
                for (int s = 0; s < KT; ++s) {
                    const float x = pv[s * 32 + lane];
                    if (!__any_sync(0xffffffffu, x >= top.v[KT - 1])) break;
                    top.insert_tie(x, pi[s * 32 + lane]);
                }
                if (qvalid) lp_finish_query<KT>(p, top, t.rg, n, q, win_lo);
            }
            thr_pub[warp * 32 + lane] = -INFINITY;
            mid_pub[warp * 32 + lane] = -INFINITY;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"((NEPI / 4) * 32) : "memory");
        }
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    if (warp == kMmaWarp) tc::tmem_dealloc_pair<512>(tmem_base);
}

// dynamic shared memory starts right after the kernel's static part; only the gap up to the next 1024-byte boundary is
// needed as alignment slack (static barriers + published thresholds + the carve-up fill the 227 KB of an SM to the byte)
template <class Kernel>
static int tc_dynamic_smem(Kernel kernel, size_t* smem_dev, size_t* smem_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    size_t smem = (dev >= 0 && dev < 64) ? smem_dev[dev] : 0;
    if (smem == 0) {     // per device; the attribute calls cost microseconds on the launch path
        cudaFuncAttributes fa;
        CRW_CUDA_RET(cudaFuncGetAttributes(&fa, kernel));
        const size_t slack = (1024 - (fa.sharedSizeBytes % 1024)) % 1024;
        smem = slack + (size_t)kQBytes + (size_t)kStages * kKBytes + (size_t)kParkBytes;
        CRW_CUDA_RET(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_dev[dev] = smem;
    }
    *smem_out = smem;
    return CRW_OK;
}

template <int KT>
static int launch_pair(const CUtensorMap* maps, const TcParams& p, int max_ctas, cudaStream_t st) {
    static size_t smem_dev[64] = {};
    size_t smem = 0;
    const int rc = tc_dynamic_smem(lp_topk_pair_kernel<KT>, smem_dev, &smem);
    if (rc != CRW_OK) return rc;
    const int items = p.v_end - p.v_begin;
    int clusters = max_ctas / 2 < 1 ? 1 : max_ctas / 2;
    if (items < clusters) clusters = items;
    lp_topk_pair_kernel<KT><<<2 * clusters, 320, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

template <int KT, int NEPI, bool TS>
static int launch_tc(const CUtensorMap* maps, const TcParams& p, int max_ctas, cudaStream_t st) {
    static size_t smem_dev[64] = {};
    size_t smem = 0;
    const int rc = tc_dynamic_smem(lp_topk_tc_kernel<KT, NEPI, TS>, smem_dev, &smem);
    if (rc != CRW_OK) return rc;
    const int items = p.v_end - p.v_begin;
    const int grid = items < max_ctas ? items : max_ctas;
    lp_topk_tc_kernel<KT, NEPI, TS><<<grid, (NEPI + 2) * 32, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
template <int KT>
static int launch_tc_any(const CUtensorMap* maps, const TcParams& p, int ts, int max_ctas, cudaStream_t st) {
    return ts ? launch_tc<KT, 8, true>(maps, p, max_ctas, st) : launch_tc<KT, 8, false>(maps, p, max_ctas, st);
}

// Host side of the tensor path, in two steps so that crw_labelprop_forward can overlap the sequential label gather with
// the bulk of the top-k: lp_tc_prepare launches the prep kernel and fills the plan (tensor maps, parameters, schedule);
// lp_tc_launch runs the top-k kernel over a range of schedule slots on at most max_ctas CTAs.
// feats [R,T,N,128] fp32 -> W, I [R,T,k,N]; scratch must hold 2 * R*T*N*128 bf16.
size_t lp_tc_split_bytes() { return (size_t)kMaxSplitTiles * 2 * 2 * 32 * kBM * sizeof(float) + (size_t)kMaxSplitTiles * 4 * sizeof(int); }

// split_ws: lp_tc_split_bytes() of scratch for the tail split (null: never split)
int lp_tc_prepare(const float* feats, int R, int T, int N, int C, int ctx, float radius, float temp, int k, int do_normalize,
                  float* W, int32_t* I, void* scratch, void* split_ws, cudaStream_t st, void* plan_storage) {
    static_assert(sizeof(LpTcPlan) <= 1024, "LpTcPlan must fit the caller's plan storage");
    LpTcPlan* plan = reinterpret_cast<LpTcPlan*>(plan_storage);
    if (C != 128 || N > 128 || N < 8 || k > 32) return CRW_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)R * T * N;
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(scratch);
    __nv_bfloat16* lo = hi + rows * 128;
    TcParams& p = plan->p;
    p.pbuf = reinterpret_cast<float*>(split_ws);
    p.pcnt = split_ws ? reinterpret_cast<int*>(reinterpret_cast<char*>(split_ws) + (size_t)kMaxSplitTiles * 2 * 2 * 32 * kBM * sizeof(float)) : nullptr;
    p.split_ok = 0;
    if (feats) {       // null: the caller fills hi / lo itself with lp_tc_prep_rows (host-streamed features)
        lp_prep_bf16_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(feats, rows, do_normalize, hi, lo, p.pcnt, p.pcnt ? kMaxSplitTiles * 4 : 0);
        CRW_LAUNCH_RET();
    }
    p.W = W; p.I = I; p.R = R; p.T = T; p.N = N; p.ctx = ctx; p.k = k;
    p.total_tiles = 0; p.v_begin = p.v_end = 0; p.tiles_per_rg = 0; p.early_per_rg = 0;
    p.hi = hi; p.lo = lo; p.total_rows = rows;
    plan->pair = 0;
    plan->ts = 0;
    if (T < 2) return CRW_OK;
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(plan->maps);
    int rc = make_tmap_bf16_k64(&maps[0], hi, (uint64_t)rows, 128, kBM);
    if (rc == CRW_OK) rc = make_tmap_bf16_k64(&maps[1], lo, (uint64_t)rows, 128, kBM);
    if (rc == CRW_OK) rc = make_tmap_bf16_k64(&maps[2], hi, (uint64_t)rows, 128, kBN);
    if (rc == CRW_OK) rc = make_tmap_bf16_k64(&maps[3], lo, (uint64_t)rows, 128, kBN);
    if (rc != CRW_OK) return rc;
    const float rc_ = ceilf(radius);
    p.rb = (rc_ - 1.0f >= (float)N) ? N : (int)rc_ - 1;
    p.inv_temp = 1.0f / temp;
    if ((uint64_t)T * N * N >= (1ull << 32)) return CRW_ERR_UNSUPPORTED;
    if ((uint64_t)T * N >= (1ull << 24)) return CRW_ERR_UNSUPPORTED;       // key rows travel as floats inside the top-k lists
    p.magic_n = (unsigned)((1ull << 32) / (unsigned)N) + 1u;
    { const char* e = getenv("CRW_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
    // The CTA-pair kernel is correct (tests run it) and has the cheaper MMA side (68 vs 78 us with the selection switched
    // off), but at BASELINE config 3 the selection epilogue bounds both kernels and the pair form pays more per-tile
    // overhead there (profiles/r01_lp_pair_anatomy.txt): opt-in until the epilogue is the smaller half.
    { const char* e = getenv("CRW_LP_PAIR"); plan->pair = (e && atoi(e) != 0) ? 1 : 0; }
    // query tile in tensor memory (TS MMAs, staged through the key ring by tcgen05.cp) is the default: MMA + TMA side 82 -> 75 us at
    // config 3, whole kernel 91.3 -> 89.5 us; CRW_LP_TS=0 selects the shared-memory query tile (SS MMAs)
    { const char* e = getenv("CRW_LP_TS"); plan->ts = ((!e || atoi(e) != 0) && !plan->pair) ? 1 : 0; }
    const int tile_rows = plan->pair ? kPairM : kBM;
    p.tiles_per_rg = ceil_div(T * N, tile_rows);
    p.total_tiles = R * p.tiles_per_rg;
    const int early = ceil_div((ctx + 2) * N, tile_rows);
    p.early_per_rg = early < p.tiles_per_rg ? early : p.tiles_per_rg;
    return CRW_OK;
}

// split_tail: the tiles of the last, partial round may be cut in two (needs the split workspace and counters zeroed by the prep
// kernel of this call, i.e. at most ONE splitting launch per lp_tc_prepare)
int lp_tc_launch(const void* plan_storage, int v_begin, int v_end, int max_ctas, cudaStream_t st, bool split_tail) {
    const LpTcPlan& plan = *reinterpret_cast<const LpTcPlan*>(plan_storage);
    if (v_end <= v_begin) return CRW_OK;
    TcParams p = plan.p;
    p.v_begin = v_begin;
    p.v_end = v_end;
    { const char* e = getenv("CRW_LP_NO_SPLIT"); p.split_ok = (split_tail && p.pbuf && !plan.pair && !(e && atoi(e))) ? 1 : 0; }
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(plan.maps);
    const int k = p.k;
    if (plan.pair) {
        if (k <= 10) return launch_pair<10>(maps, p, max_ctas, st);
        if (k <= 16) return launch_pair<16>(maps, p, max_ctas, st);
        if (k <= 20) return launch_pair<20>(maps, p, max_ctas, st);
        return launch_pair<32>(maps, p, max_ctas, st);
    }
    // (a 16-epilogue-warp instantiation, launch_tc<10, 16, false>, is supported by the kernel but measured slower: 196 vs 125 us)
    if (k <= 10) return launch_tc_any<10>(maps, p, plan.ts, max_ctas, st);
    if (k <= 16) return launch_tc_any<16>(maps, p, plan.ts, max_ctas, st);
    if (k <= 20) return launch_tc_any<20>(maps, p, plan.ts, max_ctas, st);
    return launch_tc_any<32>(maps, p, plan.ts, max_ctas, st);
}
// prep of `nrows` rows whose fp32 source sits in a staging buffer: writes hi / lo rows [row_begin, row_begin + nrows)
int lp_tc_prep_rows(const float* stage, int64_t row_begin, int64_t nrows, int64_t total_rows, int do_normalize, void* scratch,
                    cudaStream_t st) {
    if (nrows <= 0) return CRW_OK;
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(scratch);
    __nv_bfloat16* lo = hi + total_rows * 128;
    lp_prep_bf16_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(stage, nrows, do_normalize, hi + row_begin * 128, lo + row_begin * 128, nullptr, 0);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
// top-k over tiles [ta, tb) of radargram rg (at most two slot ranges: early tiles and the rest)
int lp_tc_launch_tiles(const void* plan_storage, int rg, int ta, int tb, int max_ctas, cudaStream_t st) {
    const LpTcPlan& plan = *reinterpret_cast<const LpTcPlan*>(plan_storage);
    const int E = plan.p.early_per_rg, tpr = plan.p.tiles_per_rg, R = plan.p.R;
    if (tb > tpr) tb = tpr;
    int rc = CRW_OK;
    const int ea = ta < E ? ta : E, eb = tb < E ? tb : E;
    if (eb > ea) rc = lp_tc_launch(plan_storage, rg * E + ea, rg * E + eb, max_ctas, st, false);
    const int ra = ta > E ? ta : E, rb = tb > E ? tb : E;
    if (rc == CRW_OK && rb > ra) rc = lp_tc_launch(plan_storage, R * E + rg * (tpr - E) + (ra - E), R * E + rg * (tpr - E) + (rb - E), max_ctas, st, false);
    return rc;
}
int lp_tc_tiles_per_rg(const void* plan) { return reinterpret_cast<const LpTcPlan*>(plan)->p.tiles_per_rg; }
int lp_tc_tile_rows(const void* plan) { return reinterpret_cast<const LpTcPlan*>(plan)->pair ? kPairM : kBM; }
int lp_tc_total_slots(const void* plan) { return reinterpret_cast<const LpTcPlan*>(plan)->p.total_tiles; }
int lp_tc_early_slots(const void* plan) {
    const LpTcPlan* pl = reinterpret_cast<const LpTcPlan*>(plan);
    return pl->p.R * pl->p.early_per_rg;
}

int lp_profile_read(unsigned long long* host_out, int reset) {
    if (host_out) CRW_CUDA_RET(cudaMemcpyFromSymbol(host_out, g_lp_prof, sizeof(g_lp_prof)));
    if (reset) {
        void* ptr = nullptr;
        CRW_CUDA_RET(cudaGetSymbolAddress(&ptr, g_lp_prof));
        CRW_CUDA_RET(cudaMemset(ptr, 0, sizeof(g_lp_prof)));
    }
    return CRW_OK;
}

}  // namespace crw

// test aid (host only): the work items of a top-k launch over n_tiles tiles on grid CTAs, in item order: out[3 i] = tile,
// out[3 i + 1] = half (-1 whole tile, 0 / 1 key-range half), out[3 i + 2] = index among the split tiles.  Returns the item count
// (or, with out == null, only the count).
extern "C" int crw_debug_lp_schedule(int n_tiles, int grid, int split_ok, int* out, int out_capacity_items) {
    if (n_tiles < 0 || grid < 1) return CRW_ERR_INVALID;
    const crw::Sched sc(n_tiles, grid, split_ok != 0);
    if (out) {
        if (out_capacity_items < sc.n_items) return CRW_ERR_INVALID;
        for (int i = 0; i < sc.n_items; ++i) sc.map(i, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
    }
    return sc.n_items;
}

// profiling aid: copies the per-warp phase counters (160 x 8 x 10 uint64) to host memory and optionally clears them
extern "C" int crw_debug_lp_profile(unsigned long long* host_out, int reset) { return crw::lp_profile_read(host_out, reset); }
