// walk_fused.cu -- the training walk at the reference's own sizes (N <= 64 nodes, C = 128 channels) on tcgen05: ONE launch per
// direction.
//
// Replaces src/model.py:22-46 (normalise, stride-1 affinities / tau, palindrome walk, cycle cross-entropy) and its autograd.
// Common to all kernels: frames arrive by TMA (cp.async.bulk.tensor, fp32, SWIZZLE_128B), are normalised and split into bf16
// hi / lo rows in the UMMA K-major layout; every product of the walk is a tcgen05.mma (kind::f16 on the error-compensated pairs:
// hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM) whose epilogue -- temperature, the two row softmaxes, lse - diag, softmax
// backward, F.normalize backward -- works on the accumulator row a thread owns and writes the NEXT product's operand straight back
// into shared memory.  No N x N matrix is re-laid-out or re-read by a later launch: what another CTA or the reverse pass needs
// travels as ready-made operand tiles by bulk copies (cp.async.bulk shared <-> global, L2-resident).
//   walk_fused_fwd_kernel / walk_fused_bwd_kernel              one CTA per batch element (batches that fill the SMs: B >= 96)
//   walk_fused_fwd_roles_kernel / walk_fused_bwd_roles_kernel  four CTAs per batch element, each with one ROLE, for batches that would
//        leave SMs idle (4 B <= #SMs; BASELINE config 2: B = 32).  Forward: two affinity producers (contiguous ranges of t) -> chain
//        (X_k) -> cycle (M_k, loss, G_k).  Backward: chain (Y_k) -> dA (softmax backward, the tile [T1 ; T2]) -> two dE CTAs (per
//        frame: four N = 128 products, F.normalize backward, dx).  A consumer waits for a producer's tile with an acquire spin on a
//        flag the producer releases after its bulk store completed; a CTA only ever waits for CTAs with a smaller block index, and
//        producers wait for nobody.  Only the chains are serial: one product pair and one 32-column epilogue per step.
//        Warps: 0-7 epilogues (warps w and w + 4 share accumulator rows 32 (w & 3) .., 32 columns each), 8 issues TMA / bulk loads
//        and MMAs, 9 issues the bulk stores, waits for their completion and raises the flags.
// All 128 TMEM lanes are used by stacking two 64-row problems in one tile:
//   affinity   rows = [E_a ; E_b] (the two frames in the two slots of the ring): lanes 0-63 x columns 0-63 hold E_a E_b^T,
//              lanes 64-127 x columns 64-127 hold E_b E_a^T -- A_t and A_t^T, so that BOTH softmaxes of model.py:44 are row
//              softmaxes of the thread that owns the row
//   chain      X_k = [L_k ; R_k^T]:  [L_k | .] = X_{k-1} S'_{k-1} (B operand MN-major),  [. | R_k^T] = X_{k-1} S_{k-1}^T (K-major)
//   cycle      M_k = L_k R_k = X_k[0:64] (X_k[64:128])^T, rows in lanes 0-63
// (L_k = L_{k-1} S'_{k-1}, R_k = S_{k-1} R_{k-1}, M_k = L_k R_k.)
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace crw {
namespace wf {

constexpr int kThreads = 256 + 32;                       // forward: 8 epilogue warps (warps w and w + 4 share accumulator rows 32 (w & 3) .., 32 columns each) + the issuer warp
constexpr int kRoleThreads = 256 + 64;                   // role-split kernels: 8 epilogue warps + the issuer warp + the store warp
constexpr int kBwdThreads = 128 + 32;                    // backward: 4 epilogue warps (thread = accumulator row) + the issuer warp
constexpr uint32_t kPlane128 = 128 * 128;                // [128 rows][128 B]
constexpr uint32_t kTile64 = 64 * 128;                   // [64 rows][128 B]
// ---- shared memory map of the forward kernel (bytes from a 1024-aligned base) ----
constexpr uint32_t oE = 0;                               // E planes [hi, lo][k-block 0, 1][128 rows = slot 0, slot 1][128 B]
constexpr uint32_t oX = oE + 4 * kPlane128;              // X = [L ; R^T]: hi, lo
constexpr uint32_t oSS = oX + 2 * kPlane128;             // rows of S_t: hi, lo; rows of S'_t: hi, lo
constexpr uint32_t oG = oSS + 4 * kTile64;               // rows of G_k = rowsoftmax(M_k) - I: hi, lo
constexpr uint32_t oRaw = oG + 2 * kTile64;              // raw fp32 frame: 4 channel blocks x [64 rows][128 B] (SWIZZLE_128B)
constexpr uint32_t oAst = oRaw + 4 * kTile64;            // fp32 A_t, dense [N][N] (staging for the coalesced copy-out)
constexpr uint32_t oEnd = oAst + 64 * 64 * 4;
constexpr uint32_t kSmemFwd = oEnd + 1024;

// ---- saved workspace (bytes), per batch element: invn [T][64] fp32, the T normalised frames as operand rows, then blocks j = 0 .. T-2 of operand tiles [X_j | S_j, S'_j | G_j]
// (X_0, G_0 and S_{T-2} do not exist: those parts are never written and never read) ----
constexpr size_t kSaveX = 2 * kPlane128, kSaveSS = 4 * kTile64, kSaveG = 2 * kTile64;
constexpr size_t kSaveStep = kSaveX + kSaveSS + kSaveG;                   // 80 KB
constexpr size_t kSaveFrame = 4 * kTile64;                                // the normalised frame as bf16 hi / lo operand rows: 32 KB
struct Layout {
    size_t invn, frames, tiles, part, ctr, flags, total, per_b;
    int T;
    __host__ __device__ Layout(int B, int T_) : T(T_) {
        const int K = T_ > 2 ? T_ - 2 : 0;
        invn = 0;
        frames = align_up((size_t)T_ * 64 * sizeof(float), 1024);
        tiles = frames + (size_t)T_ * kSaveFrame;
        per_b = tiles + (size_t)(K + 1) * kSaveStep;
        part = (size_t)B * per_b;
        ctr = part + align_up((size_t)2 * B * sizeof(float), 256);      // (role-split forward: chain and cycle CTA of an element each add one)
        flags = ctr + 256;
        total = flags + align_up((size_t)B * 4 * T_ * sizeof(int), 256);
    }
    // hand-over flags of the role-split kernels, per batch element [ss_ready | x_ready | y_ready | t_ready][T] (ctr + 4: error flag)
    __host__ __device__ size_t flag(int b, int which, int i) const { return flags + ((size_t)(b * 4 + which) * T + i) * sizeof(int); }
    __host__ __device__ size_t step(int b, int j) const { return (size_t)b * per_b + tiles + (size_t)j * kSaveStep; }
    __host__ __device__ size_t frame(int b, int f) const { return (size_t)b * per_b + frames + (size_t)f * kSaveFrame; }   // [hi, lo][k-block][64 rows][128 B]
};

struct FwdParams {
    int B, T, N, prof;
    int P;                    // role-split kernel: affinity producers per batch element
    float inv_tau;
    float* A;                 // [B, T-1, N, N] or null
    float* loss;              // scalar
    uint8_t* ws;              // saved workspace (1024-aligned)
};

// offset of 16-byte chunk `chunk` of row `row` inside a [rows][128 B] SWIZZLE_128B region
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(((row >> 3) << 10) + ((row & 7) << 7) + (((chunk ^ row) & 7) << 4)); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// two floats -> packed bf16 pair (hi) and the packed pair of the residuals (lo)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// one 64-element row (fp32, pads already zero) -> row `row` of a hi and a lo plane
__device__ __forceinline__ void store_row64(uint32_t hi_base, uint32_t lo_base, int row, const float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(v[8 * c + 2 * j], v[8 * c + 2 * j + 1], h[j], l[j]);
        const uint32_t o = swz(row, c);
        sts128(hi_base + o, h[0], h[1], h[2], h[3]);
        sts128(lo_base + o, l[0], l[1], l[2], l[3]);
    }
}
// row `row` of a hi / lo plane pair -> 64 floats (hi + lo)
__device__ __forceinline__ void load_row64(uint32_t hi_base, uint32_t lo_base, int row, float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t o = swz(row, c);
        const uint4 h = lds128u(hi_base + o), l = lds128u(lo_base + o);
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&hh[j]), b2 = *reinterpret_cast<const __nv_bfloat162*>(&ll[j]);
            v[8 * c + 2 * j] = __low2float(a) + __low2float(b2);
            v[8 * c + 2 * j + 1] = __high2float(a) + __high2float(b2);
        }
    }
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    tc::tmem_ld_32x32b_x32(taddr, reinterpret_cast<float(&)[32]>(v[0]));
    tc::tmem_ld_32x32b_x32(taddr + 32u, reinterpret_cast<float(&)[32]>(v[32]));
    tc::tmem_ld_wait();
}
// bulk copies shared <-> global (the TMA unit, no tensor map: whole operand tiles, layout unchanged)
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes),
                 "r"(tc::smem_u32(bar))
                 : "memory");
}

// ---- hand-over between the CTAs of the role-split kernels: operand tiles travel through global memory (L2) by bulk copies, a
// flag per tile says it is complete.  Writer (one lane): bulk stores -> wait for their completion -> fence -> release store.
// Reader (one lane): acquire load (spin) -> fence -> bulk load.  A reader only ever waits for CTAs with a SMALLER block index
// (dispatched earlier), and a writer never waits for a reader.  The spin is bounded: a lost flag ends in an error code, not a hang.
__device__ __forceinline__ void flag_set(int* f) {
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
}
__device__ __forceinline__ bool flag_peek(const int* f) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    return v != 0;
}
__device__ __forceinline__ bool flag_wait(const int* f, int* err) {
    int v = 0;
    for (long long it = 0; it < (1ll << 24); ++it) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v != 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            return true;
        }
        __nanosleep(64);
    }
    atomicExch(err, 1);
    return false;
}

// D[128 lanes x NN columns at tmem_d] (+)= A (128 rows x 64 nkb) . B (NN rows x 64 nkb)^T on bf16 hi / lo pairs: hi.hi + hi.lo + lo.hi.
// A_MN / B_MN: the operand is stored [k rows][64-element groups] instead of [rows][64 k].  *_kbs = byte distance between the
// operand's k-blocks of 64; a_lbo / b_lbo = byte distance between the 64-element groups of an MN-major operand.  One thread issues.
// (An SS MMA reads its A and B slices from shared memory: at N = 64 that is 6 KB per 33 cycles of tensor time, more than the
// 128 B / cycle the SM delivers -- measured ~55 cycles per MMA; N = 128 amortises the A slice.)
template <bool A_MN, bool B_MN, int NN = 64>
__device__ __forceinline__ void mma3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int nkb, uint32_t a_kbs,
                                     uint32_t b_kbs, uint32_t a_lbo, bool fresh, uint32_t b_lbo = 8192) {
    const uint32_t idesc = tc::umma_idesc_bf16_major(128, NN, A_MN, B_MN);
    bool first = fresh;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t ap = (pass == 2) ? a_lo : a_hi, bp = (pass == 1) ? b_lo : b_hi;
        for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = A_MN ? tc::umma_smem_desc_mn128(ap + kb * a_kbs + ks * 2048, a_lbo, 1024) : tc::umma_smem_desc_k128(ap + kb * a_kbs + ks * 32);
                const uint64_t bd = B_MN ? tc::umma_smem_desc_mn128(bp + kb * b_kbs + ks * 2048, b_lbo, 1024) : tc::umma_smem_desc_k128(bp + kb * b_kbs + ks * 32);
                tc::umma_bf16_ss(tmem_d, ad, bd, idesc, first ? 0u : 1u);
                first = false;
            }
        }
    }
}
// 32 columns of a row (fp32, pads already zero) -> chunks 4 ch .. 4 ch + 3 of row `row` of a hi and a lo plane
__device__ __forceinline__ void store_row32(uint32_t hi_base, uint32_t lo_base, int row, int ch, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(v[8 * c + 2 * j], v[8 * c + 2 * j + 1], h[j], l[j]);
        const uint32_t o = swz(row, 4 * ch + c);
        sts128(hi_base + o, h[0], h[1], h[2], h[3]);
        sts128(lo_base + o, l[0], l[1], l[2], l[3]);
    }
}
__device__ __forceinline__ void load_row32(uint32_t hi_base, uint32_t lo_base, int row, int ch, float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t o = swz(row, 4 * ch + c);
        const uint4 h = lds128u(hi_base + o), l = lds128u(lo_base + o);
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&hh[j]), b2 = *reinterpret_cast<const __nv_bfloat162*>(&ll[j]);
            v[8 * c + 2 * j] = __low2float(a) + __low2float(b2);
            v[8 * c + 2 * j + 1] = __high2float(a) + __high2float(b2);
        }
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    tc::tmem_ld_32x32b_x32(taddr, v);
    tc::tmem_ld_wait();
}

// profiling aid (CRW_WALK_PROF=1): cycles per phase of CTA 0 as seen by thread 0, [kernel 0 = forward, 1 = backward][16 phases]
__device__ unsigned long long g_wf_prof[2 * 16];
#define WFPROF(kern, idx) do { if (p.prof && blockIdx.x == 0 && tid == 0) { const long long c_now = clock64(); g_wf_prof[(kern) * 16 + (idx)] += (unsigned long long)(c_now - c_last); c_last = c_now; } } while (0)

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) walk_fused_fwd_kernel(const __grid_constant__ CUtensorMap xmap, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_red[4][64];           // row reductions across the threads that share a row
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x, N = p.N, T = p.T, K = T - 2;
    const Layout lay(p.B, T);

    if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_tma, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
        tc::prefetch_tmap(&xmap);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // Warps 0-7 own the accumulator rows (epilogues): warp w works on TMEM lanes 32 (w & 3) .. and on column half w >> 2 of its
    // row's 64-column window.  Warp 8 issues the TMA loads and the MMAs from ONE lane elected inside its own branch, so that the
    // compiler emits straight-line issue code (inside an `if (tid == 0)` region every tcgen05.mma is wrapped in an elect + branch
    // loop).  Both roles walk the same sequence of CTA barriers.
    const bool is_epi = warp < 8, is_iss = warp == 8;
    const uint32_t tmem = tmem_base_s;
    const uint32_t tA = tmem, tX = tmem + 128u, tM = tmem + 256u;       // accumulators: affinity (128 cols), chain (128), cycle (64)
    const int row = ((warp & 3) << 5) | lane;                            // accumulator row (TMEM lane)
    const int half = row >> 6, r = row & 63, ch = (warp >> 2) & 1;       // which 64-row problem, row in it, column half
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t tma_phase = 0, mma_phase = 0;

    const uint32_t sE = sb + oE, sX = sb + oX, sSS = sb + oSS, sG = sb + oG, sRaw = sb + oRaw, sAst = sb + oAst;
    auto ePlane = [&](int pl, int kb) { return sE + (uint32_t)(pl * 2 + kb) * kPlane128; };
    auto publish = [&]() {                   // operand tiles written / accumulators read -> visible to the MMAs issued after the barrier
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
    };
    auto mma_wait = [&]() {
        tc::mbar_wait(&bar_mma, mma_phase & 1);
        ++mma_phase;
        tc::tc_fence_after();
    };

    auto issue_frame = [&](int f) {          // issuer lane: frame f of this batch element -> raw staging (four 32-channel blocks)
        tc::mbar_arrive_expect_tx(&bar_tma, (uint32_t)N * 512u);
        const int row0 = (b * T + f) * N;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb)
            tc::tma_load_2d(reinterpret_cast<void*>(smem_raw + (sRaw - tc::smem_u32(smem_raw)) + cb * kTile64), &xmap, cb * 32, row0, &bar_tma);
    };
    // raw staging -> normalised bf16 hi / lo rows of ring slot f & 1; thread = (32-channel block cq, row rr).  Holds one CTA barrier.
    auto convert_frame = [&](int f) {
        const int cq = (tid >> 6) & 3, rr = tid & 63;
        float v[32];
        float ss = 0.0f;
        if (is_epi && rr < N) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = lds128f(sRaw + (uint32_t)cq * kTile64 + swz(rr, c));
                v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
                ss = fmaf(q.x, q.x, ss); ss = fmaf(q.y, q.y, ss); ss = fmaf(q.z, q.z, ss); ss = fmaf(q.w, q.w, ss);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
        }
        if (is_epi) s_red[cq][rr] = ss;
        __syncthreads();
        if (is_epi) {
            const float inv = 1.0f / fmaxf(sqrtf((s_red[0][rr] + s_red[1][rr]) + (s_red[2][rr] + s_red[3][rr])), kNormEps);
            if (cq == 0) reinterpret_cast<float*>(p.ws + (size_t)b * lay.per_b + lay.invn)[f * 64 + rr] = (rr < N) ? inv : 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv;
            store_row32(ePlane(0, cq >> 1), ePlane(1, cq >> 1), (f & 1) * 64 + rr, cq & 1, v);
        }
    };
    auto save_frame = [&](int f) {           // issuer lane: the operand rows of frame f, for the reverse pass
        uint8_t* dst = p.ws + lay.frame(b, f);
#pragma unroll
        for (int q = 0; q < 4; ++q) bulk_store(dst + q * kTile64, sE + (uint32_t)q * kPlane128 + (uint32_t)(f & 1) * kTile64, kTile64);
        bulk_commit();
    };

    if (is_iss && tc::elect_one()) issue_frame(0);
    tc::mbar_wait(&bar_tma, tma_phase & 1); ++tma_phase;
    convert_frame(0);
    publish();                                                           // raw staging is free again
    if (is_iss && tc::elect_one()) { issue_frame(1); save_frame(0); }

    float loss_acc = 0.0f;
    long long c_last = clock64();
    for (int t = 0; t + 1 < T; ++t) {
        if (t >= K && !p.A) break;                                       // the last affinity only feeds the returned A
        tc::mbar_wait(&bar_tma, tma_phase & 1); ++tma_phase;
        WFPROF(0, 0);
        convert_frame(t + 1);
        WFPROF(0, 1);
        publish();
        if (is_iss && tc::elect_one()) {
            tc::tc_fence_after();
            if (t + 2 < T && (t + 1 < K || p.A)) issue_frame(t + 2);
            save_frame(t + 1);
            // affinity: [rows . slot 0 | rows . slot 1] in one N = 128 product
            mma3<false, false, 128>(tA, ePlane(0, 0), ePlane(1, 0), ePlane(0, 0), ePlane(1, 0), 2, kPlane128, kPlane128, 0, true);
            // everything saved before this step's frame has left shared memory before the epilogues below overwrite S / S' / X / G
            // (and, one step on, the other ring slot): the wait runs under the MMAs
            bulk_wait_read1();
            tc::umma_commit(&bar_mma);
        }
        WFPROF(0, 2);
        mma_wait();
        WFPROF(0, 3);
        // lanes 0-63 are the rows of the frame in slot 0 and want the product with slot 1 (columns 64-127), and the other way round
        const bool isA = (half == 0) == ((t & 1) == 0);                  // this thread's accumulator row is a row of A_t (else of A_t^T)
        float a[32];
        float mx = -INFINITY;
        if (is_epi) {
            tmem_ld32(tA + lane_base + (uint32_t)((1 - half) * 64 + ch * 32), a);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                a[c] *= p.inv_tau;
                if (32 * ch + c < N) mx = fmaxf(mx, a[c]);
            }
            if (p.A && isA && r < N) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (32 * ch + c < N) tc::sts_f32(sAst + (uint32_t)(r * N + 32 * ch + c) * 4u, a[c]);
            }
            s_red[half * 2 + ch][r] = mx;
        }
        tc::tc_fence_before();
        __syncthreads();
        float s = 0.0f;
        if (is_epi) {
            mx = fmaxf(mx, s_red[half * 2 + (1 - ch)][r]);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                a[c] = (32 * ch + c < N && r < N) ? exp2f((a[c] - mx) * 1.4426950408889634f) : 0.0f;
                s += a[c];
            }
        }
        __syncthreads();
        if (is_epi) s_red[half * 2 + ch][r] = s;
        __syncthreads();
        if (is_epi) {
            s += s_red[half * 2 + (1 - ch)][r];
            const float is = (r < N) ? 1.0f / s : 0.0f;
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] *= is;
            if (t < K) {
                // rows of S_t -> the K-major B operand of [. | R^T]; rows of S'_t -> the MN-major B operand of [L | .]
                store_row32(sSS + (isA ? 0u : 2u * kTile64), sSS + (isA ? 1u : 3u) * kTile64, r, ch, a);
                if (t == 0) {
                    // X_1 = [L_1 ; R_1^T] = [S'_0 ; I]
                    if (!isA) {
                        store_row32(sX, sX + kPlane128, r, ch, a);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) a[c] = (32 * ch + c == r && r < N) ? 1.0f : 0.0f;
                        store_row32(sX, sX + kPlane128, 64 + r, ch, a);
                    }
                }
            }
        }
        publish();
        WFPROF(0, 4);
        if (p.A) {
            float* dst = p.A + ((size_t)b * (T - 1) + t) * N * N;
            for (int i = tid; i < N * N; i += kThreads) dst[i] = tc::lds_f32(sAst + (uint32_t)i * 4u);
        }
        if (t >= K) continue;
        const int k = t + 1;
        if (k >= 2) {
            if (is_iss && tc::elect_one()) {
                tc::tc_fence_after();
                mma3<false, true>(tX, sX, sX + kPlane128, sSS + 2 * kTile64, sSS + 3 * kTile64, 1, 0, 0, 0, true);      // X . S'_{k-1}
                mma3<false, false>(tX + 64u, sX, sX + kPlane128, sSS, sSS + kTile64, 1, 0, 0, 0, true);                  // X . S_{k-1}^T
                tc::umma_commit(&bar_mma);
            }
            WFPROF(0, 5);
            mma_wait();
            WFPROF(0, 6);
            if (is_epi) {
                float x[32];
                tmem_ld32(tX + lane_base + (uint32_t)(half * 64 + ch * 32), x);
                store_row32(sX, sX + kPlane128, row, ch, x);
            }
            publish();
            WFPROF(0, 7);
        }
        if (is_iss && tc::elect_one()) {
            tc::tc_fence_after();
            mma3<false, false>(tM, sX, sX + kPlane128, sX + kTile64, sX + kPlane128 + kTile64, 1, 0, 0, 0, true);        // M_k = L_k R_k
            tc::umma_commit(&bar_mma);
        }
        WFPROF(0, 8);
        mma_wait();
        WFPROF(0, 9);
        // rows of M_k (lanes 0-63): lse - diag and G_k = rowsoftmax(M_k) - I.  The entries of M_k lie in [0, 1]: exp(m - 1) needs no row max
        float m[32];
        float sm = 0.0f, diag = 0.0f;
        if (is_epi && half == 0) {
            tmem_ld32(tM + lane_base + (uint32_t)(ch * 32), m);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                if (32 * ch + c == r) diag = m[c];
                m[c] = (32 * ch + c < N && r < N) ? exp2f((m[c] - 1.0f) * 1.4426950408889634f) : 0.0f;
                sm += m[c];
            }
            s_red[ch][r] = sm;
        }
        tc::tc_fence_before();
        __syncthreads();
        if (is_epi && half == 0) {
            sm += s_red[1 - ch][r];
            if (r < N && (r >> 5) == ch) loss_acc += logf(sm) + 1.0f - diag;
            const float is = (r < N) ? 1.0f / sm : 0.0f;
#pragma unroll
            for (int c = 0; c < 32; ++c) m[c] = m[c] * is - ((32 * ch + c == r && r < N) ? 1.0f : 0.0f);
            store_row32(sG, sG + kTile64, r, ch, m);
        }
        publish();
        if (is_iss && tc::elect_one()) {     // what the reverse pass needs of step k, as ready-made operand tiles
            uint8_t* dst = p.ws + lay.step(b, k);
            bulk_store(dst, sX, (uint32_t)kSaveX);
            bulk_store(p.ws + lay.step(b, k - 1) + kSaveX, sSS, (uint32_t)kSaveSS);
            bulk_store(dst + kSaveX + kSaveSS, sG, (uint32_t)kSaveG);
            bulk_commit();
        }
        WFPROF(0, 10);
    }
    // loss = sum over (b, k, d) of (lse - diag) / (B N N); the last CTA to finish adds the per-element sums in order
    __syncthreads();
    if (is_epi && half == 0) s_red[ch][r] = loss_acc;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.0f;
        for (int i = 0; i < 64; ++i) tot += s_red[0][i] + s_red[1][i];
        float* part = reinterpret_cast<float*>(p.ws + lay.part);
        part[b] = tot;
        __threadfence();
        const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(p.ws + lay.ctr), 1u);
        if (done == (unsigned)p.B - 1u) {
            __threadfence();
            float sum = 0.0f;
            for (int i = 0; i < p.B; ++i) sum += reinterpret_cast<volatile float*>(part)[i];
            *p.loss = sum / ((float)p.B * (float)N * (float)N);
        }
    }
    if (is_iss && tc::elect_one()) bulk_wait_all();                      // (bulk groups belong to the thread that issued them)
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// forward, role-split: P + 2 CTAs per batch element (grid = (P + 2) B, role = blockIdx.x / B), for batches that leave SMs idle.
//   roles < P   producers: affinities t = role, role + P, ...: frames by TMA -> normalise -> one N = 128 product -> both softmaxes ->
//               rows of S_t / S'_t as operand tiles -> global (block t of the saved workspace, where the reverse pass wants them
//               anyway) + flag; also A_t, invn and the frames' operand rows for the reverse pass
//   role P      chain: X_1 = [S'_0 ; I], X_k = X_{k-1} [S'_{k-1} | S_{k-1}^T]: waits for S tiles, saves X_k + flag
//   role P + 1  cycle: M_k = L_k R_k from the saved X_k: lse - diag, G_k, the loss
// The only serial part left on one SM is the chain (one product + one 32-column epilogue per step).
// ------------------------------------------------------------------------------------------
constexpr uint32_t rE = 0;                               // producer: E planes (64 KB) | chain: S tiles, three buffers, then X | cycle: X, two buffers
constexpr uint32_t rRaw = rE + 4 * kPlane128;            // producer: raw frames, two buffers (64 KB) | cycle: G (16 KB)
constexpr uint32_t rSS = rRaw + 8 * kTile64;             // producer: S tiles out (32 KB)
constexpr uint32_t rAst = rSS + 4 * kTile64;             // producer: fp32 A_t staging (16 KB)
constexpr uint32_t rEnd = rAst + 64 * 64 * 4;
constexpr uint32_t kSmemRoles = rEnd + 1024;

__global__ void __launch_bounds__(kRoleThreads, 1) walk_fused_fwd_roles_kernel(const __grid_constant__ CUtensorMap xmap, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sb - tc::smem_u32(smem_raw));            // generic pointer to the aligned base
    __shared__ uint64_t bar_ld[2], bar_mma, bar_st, bar_tile, bar_ld3[3];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_red[4][64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = blockIdx.x / p.B, b = blockIdx.x % p.B, N = p.N, T = p.T, K = T - 2;
    const Layout lay(p.B, T);
    int* err = reinterpret_cast<int*>(p.ws + lay.ctr + 4);
    auto flagp = [&](int which, int i) { return reinterpret_cast<int*>(p.ws + lay.flag(b, which, i)); };

    if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_ld[0], 1);
        tc::mbar_init(&bar_ld[1], 1);
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_init(&bar_st, 1);
        tc::mbar_init(&bar_tile, 1);
        for (int i = 0; i < 3; ++i) tc::mbar_init(&bar_ld3[i], 1);
        tc::fence_barrier_init();
        tc::prefetch_tmap(&xmap);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // warp 8 issues loads and MMAs; warp 9 issues the bulk stores, waits for their completion and raises the flags -- it is back at the
    // next CTA barrier before anybody overwrites what it copied, and its waits stay off the issuer's path
    const bool is_epi = warp < 8, is_iss = warp == 8, is_st = warp == 9;
    const uint32_t tmem = tmem_base_s;
    const int row = ((warp & 3) << 5) | lane;
    const int half = row >> 6, r = row & 63, ch = (warp >> 2) & 1;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ld_phase[2] = {0u, 0u}, mma_phase = 0;
    // the store warp's copy of a tile has LEFT shared memory (arrives on bar_st): waited for by everybody right before the tile is
    // overwritten in place; its completion in global memory and the flag follow on the store lane alone
    uint32_t st_phase = 0;
    auto st_wait = [&]() {
        if (is_st) return;                   // (the store warp is the one that arrives; see ld_wait)
        tc::mbar_wait(&bar_st, st_phase & 1);
        ++st_phase;
    };
    auto store_tile = [&](void* gdst, uint32_t ssrc, uint32_t bytes, int* flag) {     // store lane
        bulk_store(gdst, ssrc, bytes);
        bulk_commit();
        bulk_wait_read();
        tc::mbar_arrive(&bar_st);
        bulk_wait_all();
        if (flag) flag_set(flag);
    };
    auto publish = [&]() {
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
    };
    auto mma_wait_raw = [&]() {
        tc::mbar_wait(&bar_mma, mma_phase & 1);
        ++mma_phase;
        tc::tc_fence_after();
    };
    // profiling aid (CRW_WALK_PROF=1): per role of element 0, cycles [total, waiting for loads (hand-over + TMA), waiting for MMAs]
    const bool rprof = p.prof && b == 0 && tid == 0;
    const long long c_begin = clock64();
    long long c_ld = 0, c_mma = 0;
    // (The store warp skips the waits on load / MMA barriers: it needs none of that data, and, coming late from a copy it completed,
    // it could find a barrier re-armed and completed AGAIN -- same parity -- and wait for ever.)
    auto ld_wait = [&](int i) {
        if (is_st) return;
        const long long c0 = rprof ? clock64() : 0;
        tc::mbar_wait(&bar_ld[i], ld_phase[i] & 1);
        ++ld_phase[i];
        if (rprof) c_ld += clock64() - c0;
    };
    auto mma_wait = [&]() {
        if (is_st) return;
        const long long c0 = rprof ? clock64() : 0;
        mma_wait_raw();
        if (rprof) c_mma += clock64() - c0;
    };

    // loss = sum over (b, k, d) of (lse - diag) / (B N N): the chain CTA (last cycle step) and the cycle CTA of every element each
    // deposit one partial sum; the last of the 2 B depositors adds them in order
    auto deposit_loss = [&](float tot, int slot) {       // one thread
        float* part = reinterpret_cast<float*>(p.ws + lay.part);
        part[slot] = tot;
        __threadfence();
        const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(p.ws + lay.ctr), 1u);
        if (done == 2u * (unsigned)p.B - 1u) {
            __threadfence();
            float sum = 0.0f;
            for (int i = 0; i < 2 * p.B; ++i) sum += reinterpret_cast<volatile float*>(part)[i];
            if (*reinterpret_cast<volatile int*>(err)) sum = __int_as_float(0x7fc00000);      // a hand-over flag never arrived: NaN, not a wrong number
            *p.loss = sum / ((float)p.B * (float)N * (float)N);
        }
    };
    const int P = p.P;
    if (role < P) {
        // ================= producer =================
        const int Tlast = p.A ? T - 1 : K;                               // affinities 0 .. Tlast - 1 are wanted
        const uint32_t sE = sb + rE, sSS = sb + rSS, sAst = sb + rAst;
        auto ePlane = [&](int pl, int kb) { return sE + (uint32_t)(pl * 2 + kb) * kPlane128; };
        auto issue_frame = [&](int f, int buf) {     // issuer lane
            tc::mbar_arrive_expect_tx(&bar_ld[buf], (uint32_t)N * 512u);
            const int row0 = (b * T + f) * N;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) tc::tma_load_2d(sgen + rRaw + (buf * 4 + cb) * kTile64, &xmap, cb * 32, row0, &bar_ld[buf]);
        };
        auto convert_frame = [&](int f, int buf) {   // raw buffer `buf` -> ring slot f & 1; one CTA barrier inside
            const int cq = (tid >> 6) & 3, rr = tid & 63;
            float v[32];
            float ss = 0.0f;
            if (is_epi && rr < N) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 q = lds128f(sb + rRaw + (uint32_t)(buf * 4 + cq) * kTile64 + swz(rr, c));
                    v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
                    ss = fmaf(q.x, q.x, ss); ss = fmaf(q.y, q.y, ss); ss = fmaf(q.z, q.z, ss); ss = fmaf(q.w, q.w, ss);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (is_epi) s_red[cq][rr] = ss;
            __syncthreads();
            if (is_epi) {
                const float inv = 1.0f / fmaxf(sqrtf((s_red[0][rr] + s_red[1][rr]) + (s_red[2][rr] + s_red[3][rr])), kNormEps);
                if (cq == 0) reinterpret_cast<float*>(p.ws + (size_t)b * lay.per_b + lay.invn)[f * 64 + rr] = (rr < N) ? inv : 0.0f;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= inv;
                store_row32(ePlane(0, cq >> 1), ePlane(1, cq >> 1), (f & 1) * 64 + rr, cq & 1, v);
            }
        };
        auto save_frame = [&](int f) {               // issuer lane
            uint8_t* dst = p.ws + lay.frame(b, f);
#pragma unroll
            for (int q = 0; q < 4; ++q) bulk_store(dst + q * kTile64, sE + (uint32_t)q * kPlane128 + (uint32_t)(f & 1) * kTile64, kTile64);
        };
        // a CONTIGUOUS range of affinities per producer: consecutive affinities share a frame, so only the range's first one converts two
        // frames.  Raw buffer = frame parity; the next frame is fetched by TMA while this affinity's product and epilogue run.
        const int per = (Tlast + P - 1) / P, t0 = role * per, t1 = min(Tlast, t0 + per);
        if (t0 < t1 && is_iss) { if (tc::elect_one()) { issue_frame(t0, t0 & 1); issue_frame(t0 + 1, (t0 + 1) & 1); } __syncwarp(); }
        for (int t = t0; t < t1; ++t) {
            if (t == t0) {
                ld_wait(t & 1);
                convert_frame(t, t & 1);
            }
            ld_wait((t + 1) & 1);
            convert_frame(t + 1, (t + 1) & 1);
            publish();
            if (is_st) { if (tc::elect_one()) {          // (waited for together with the S tiles at the end of this affinity, before the ring is rewritten)
                if (t == t0) save_frame(t);
                save_frame(t + 1);
                bulk_commit();
            } __syncwarp(); }
            if (is_iss) { if (tc::elect_one()) {
                tc::tc_fence_after();
                if (t + 1 < t1) issue_frame(t + 2, t & 1);               // (raw buffer t & 1 held frame t: converted)
                mma3<false, false, 128>(tmem, ePlane(0, 0), ePlane(1, 0), ePlane(0, 0), ePlane(1, 0), 2, kPlane128, kPlane128, 0, true);
                tc::umma_commit(&bar_mma);
            } __syncwarp(); }
            mma_wait();
            const bool isA = (half == 0) == ((t & 1) == 0);
            float a[32];
            float mx = -INFINITY;
            if (is_epi) {
                tmem_ld32(tmem + lane_base + (uint32_t)((1 - half) * 64 + ch * 32), a);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    a[c] *= p.inv_tau;
                    if (32 * ch + c < N) mx = fmaxf(mx, a[c]);
                }
                if (p.A && isA && r < N) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (32 * ch + c < N) tc::sts_f32(sAst + (uint32_t)(r * N + 32 * ch + c) * 4u, a[c]);
                }
                s_red[half * 2 + ch][r] = mx;
            }
            tc::tc_fence_before();
            __syncthreads();
            float s = 0.0f;
            if (is_epi) {
                mx = fmaxf(mx, s_red[half * 2 + (1 - ch)][r]);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    a[c] = (32 * ch + c < N && r < N) ? exp2f((a[c] - mx) * 1.4426950408889634f) : 0.0f;
                    s += a[c];
                }
            }
            __syncthreads();
            if (is_epi) s_red[half * 2 + ch][r] = s;
            __syncthreads();
            if (is_epi && t < K) {
                s += s_red[half * 2 + (1 - ch)][r];
                const float is = (r < N) ? 1.0f / s : 0.0f;
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] *= is;
                store_row32(sSS + (isA ? 0u : 2u * kTile64), sSS + (isA ? 1u : 3u) * kTile64, r, ch, a);
            }
            publish();
            if (is_st) { if (tc::elect_one()) {
                if (t < K) {
                    bulk_store(p.ws + lay.step(b, t) + kSaveX, sSS, (uint32_t)kSaveSS);
                    bulk_commit();
                }
                bulk_wait_all();
                if (t < K) flag_set(flagp(0, t));
            } __syncwarp(); }
            if (p.A) {
                float* dst = p.A + ((size_t)b * (T - 1) + t) * N * N;
                for (int i = tid; i < N * N; i += kRoleThreads) dst[i] = tc::lds_f32(sAst + (uint32_t)i * 4u);
            }
        }
    } else if (role == P) {
        // ================= chain =================
        // The serial role: per step one product pair and one 32-column epilogue.  Everything else is kept off its path: the S tiles are
        // fetched up to two steps ahead (three buffers, only when already flagged), and the store warp runs on its own -- it waits for
        // "X_k is written" (bar_tile), copies, says "the copy has left shared memory" (bar_st), completes, flags; the other nine warps
        // synchronise among themselves on a named barrier.
        const uint32_t sSSb = sb + rE, sX = sb + rE + 3 * (uint32_t)kSaveSS;      // S tiles: three 32 KB buffers; X: hi, lo
        const uint32_t tX = tmem;
        const uint32_t sGc = sSSb + (uint32_t)(K % 3) * (uint32_t)kSaveSS;       // G_K: the S buffer the last product did not read
        if (is_st) {
            for (int k = 1; k <= K; ++k) {
                tc::mbar_wait(&bar_tile, (k - 1) & 1);
                if (tc::elect_one()) store_tile(p.ws + lay.step(b, k), sX, (uint32_t)kSaveX, flagp(1, k));      // the cycle CTA is waiting for X_k
                __syncwarp();
            }
            tc::mbar_wait(&bar_tile, K & 1);
            if (tc::elect_one()) store_tile(p.ws + lay.step(b, K) + kSaveX + kSaveSS, sGc, (uint32_t)kSaveG, nullptr);
            __syncwarp();
        } else {
            auto sync9 = [&]() {                     // operand tiles written / accumulators read -> visible; the nine working warps only
                tc::fence_proxy_async();
                tc::tc_fence_before();
                asm volatile("bar.sync 1, 288;" ::: "memory");
            };
            int next_load = 0;                       // S tiles 0 .. next_load - 1 are in flight or loaded (issuer lane's view)
            auto try_loads = [&](int upto, bool must) {      // issuer lane: fetch S_t for t < upto in order; `must`: wait for the flags
                while (next_load < upto && next_load < K) {
                    if (!must && !flag_peek(flagp(0, next_load))) break;
                    flag_wait(flagp(0, next_load), err);
                    const int buf = next_load % 3;
                    tc::mbar_arrive_expect_tx(&bar_ld3[buf], (uint32_t)kSaveSS);
                    bulk_load(sSSb + (uint32_t)buf * (uint32_t)kSaveSS, p.ws + lay.step(b, next_load) + kSaveX, (uint32_t)kSaveSS, &bar_ld3[buf]);
                    ++next_load;
                }
            };
            uint32_t ld3_phase[3] = {0u, 0u, 0u};
            auto ld3_wait = [&](int buf) {
                tc::mbar_wait(&bar_ld3[buf], ld3_phase[buf] & 1);
                ++ld3_phase[buf];
            };
            if (is_iss) { if (tc::elect_one()) { try_loads(1, true); try_loads(3, false); } __syncwarp(); }
            for (int k = 1; k <= K; ++k) {
                const int buf = (k - 1) % 3;
                const uint32_t sSS = sSSb + (uint32_t)buf * (uint32_t)kSaveSS;
                if (is_iss) { if (tc::elect_one()) try_loads(k, true); __syncwarp(); }      // (S_{k-1}: waited for here, where nothing else is held up)
                ld3_wait(buf);
                if (k == 1) {
                    // X_1 = [L_1 ; R_1^T] = [S'_0 ; I]: the rows of S'_0 are already operand rows (same swizzle): copy them
                    if (is_epi) {
                        for (int i = tid; i < 2 * 512; i += 256) {           // 2 planes x 512 chunks of 16 bytes
                            const int pl = i >> 9, cidx = i & 511;
                            const uint4 q = lds128u(sSS + (uint32_t)(2 + pl) * kTile64 + (uint32_t)cidx * 16u);
                            sts128(sX + (uint32_t)pl * kPlane128 + (uint32_t)cidx * 16u, q.x, q.y, q.z, q.w);
                        }
                        if (half == 1) {
                            float e[32];
#pragma unroll
                            for (int c = 0; c < 32; ++c) e[c] = (32 * ch + c == r && r < N) ? 1.0f : 0.0f;
                            store_row32(sX, sX + kPlane128, 64 + r, ch, e);
                        }
                    }
                    sync9();
                    if (is_iss) { if (tc::elect_one()) { tc::mbar_arrive(&bar_tile); try_loads(4, false); } __syncwarp(); }
                } else {
                    if (is_iss) { if (tc::elect_one()) {
                        tc::tc_fence_after();
                        mma3<false, true>(tX, sX, sX + kPlane128, sSS + 2 * kTile64, sSS + 3 * kTile64, 1, 0, 0, 0, true);   // X . S'_{k-1}
                        mma3<false, false>(tX + 64u, sX, sX + kPlane128, sSS, sSS + kTile64, 1, 0, 0, 0, true);               // X . S_{k-1}^T
                        tc::umma_commit(&bar_mma);
                        try_loads(k + 2, false);         // under the MMAs: later S tiles whose producers are already done (never the
                                                         // buffer these MMAs read); no waiting here -- the epilogue needs this warp
                    } __syncwarp(); }
                    mma_wait();
                    st_wait();                                               // the copy of X_{k-1} has left shared memory
                    if (is_epi) {
                        float x[32];
                        tmem_ld32(tX + lane_base + (uint32_t)(half * 64 + ch * 32), x);
                        store_row32(sX, sX + kPlane128, row, ch, x);
                    }
                    sync9();
                    if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
                }
            }
            // the LAST cycle step here, where X_K already is: M_K = L_K R_K, lse - diag, G_K (the cycle CTA does steps 1 .. K - 1; handing
            // X_K over would put a flag, a 32 KB load and a product latency at the very end of the launch)
            const uint32_t tM = tmem + 128u;
            if (is_iss) { if (tc::elect_one()) {
                tc::tc_fence_after();
                mma3<false, false>(tM, sX, sX + kPlane128, sX + kTile64, sX + kPlane128 + kTile64, 1, 0, 0, 0, true);
                tc::umma_commit(&bar_mma);
            } __syncwarp(); }
            mma_wait();
            float m[32];
            float sm = 0.0f, diag = 0.0f, lsum = 0.0f;
            if (is_epi && half == 0) {
                tmem_ld32(tM + lane_base + (uint32_t)(ch * 32), m);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (32 * ch + c == r) diag = m[c];
                    m[c] = (32 * ch + c < N && r < N) ? exp2f((m[c] - 1.0f) * 1.4426950408889634f) : 0.0f;
                    sm += m[c];
                }
                s_red[ch][r] = sm;
            }
            tc::tc_fence_before();
            asm volatile("bar.sync 1, 288;" ::: "memory");
            if (is_epi && half == 0) {
                sm += s_red[1 - ch][r];
                if (r < N && (r >> 5) == ch) lsum = logf(sm) + 1.0f - diag;
                const float is = (r < N) ? 1.0f / sm : 0.0f;
#pragma unroll
                for (int c = 0; c < 32; ++c) m[c] = m[c] * is - ((32 * ch + c == r && r < N) ? 1.0f : 0.0f);
                store_row32(sGc, sGc + kTile64, r, ch, m);
            }
            sync9();
            if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
            if (is_epi && half == 0) s_red[2 + ch][r] = lsum;
            asm volatile("bar.sync 1, 288;" ::: "memory");
            if (tid == 0) {
                float tot = 0.0f;
                for (int i = 0; i < 64; ++i) tot += s_red[2][i] + s_red[3][i];
                deposit_loss(tot, p.B + b);
            }
        }
    } else {
        // ================= cycle =================
        const uint32_t sXb = sb + rE, sG = sb + rRaw;                    // X: two 32 KB buffers; G: hi, lo
        const uint32_t tM = tmem;
        auto load_x = [&](int k, int buf) {          // issuer lane
            flag_wait(flagp(1, k), err);
            tc::mbar_arrive_expect_tx(&bar_ld[buf], (uint32_t)kSaveX);
            bulk_load(sXb + (uint32_t)buf * (uint32_t)kSaveX, p.ws + lay.step(b, k), (uint32_t)kSaveX, &bar_ld[buf]);
        };
        float loss_acc = 0.0f;
        int x_issued = 0;                            // (issuer lane) X_k has been fetched for k <= x_issued
        for (int k = 1; k < K; ++k) {                // (the chain CTA does step K itself)
            const int buf = k & 1;
            const uint32_t sX = sXb + (uint32_t)buf * (uint32_t)kSaveX;
            if (is_iss) { if (tc::elect_one()) { if (x_issued < k) { load_x(k, buf); x_issued = k; } } __syncwarp(); }      // (waits for the flag here)
            ld_wait(buf);
            if (is_iss) { if (tc::elect_one()) {
                tc::tc_fence_after();
                mma3<false, false>(tM, sX, sX + kPlane128, sX + kTile64, sX + kPlane128 + kTile64, 1, 0, 0, 0, true);    // M_k = L_k R_k
                tc::umma_commit(&bar_mma);
                // the next X under this step's epilogue -- only if it is already there: the epilogue needs this warp at its barriers
                if (k + 1 < K && flag_peek(flagp(1, k + 1))) { load_x(k + 1, (k + 1) & 1); x_issued = k + 1; }
            } __syncwarp(); }
            mma_wait();
            if (k >= 2) st_wait();                                       // the copy of G_{k-1} has left shared memory
            float m[32];
            float sm = 0.0f, diag = 0.0f;
            if (is_epi && half == 0) {
                tmem_ld32(tM + lane_base + (uint32_t)(ch * 32), m);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (32 * ch + c == r) diag = m[c];
                    m[c] = (32 * ch + c < N && r < N) ? exp2f((m[c] - 1.0f) * 1.4426950408889634f) : 0.0f;
                    sm += m[c];
                }
                s_red[ch][r] = sm;
            }
            tc::tc_fence_before();
            __syncthreads();
            if (is_epi && half == 0) {
                sm += s_red[1 - ch][r];
                if (r < N && (r >> 5) == ch) loss_acc += logf(sm) + 1.0f - diag;
                const float is = (r < N) ? 1.0f / sm : 0.0f;
#pragma unroll
                for (int c = 0; c < 32; ++c) m[c] = m[c] * is - ((32 * ch + c == r && r < N) ? 1.0f : 0.0f);
                store_row32(sG, sG + kTile64, r, ch, m);
            }
            publish();
            if (is_st) { if (tc::elect_one()) {      // nobody waits for G_k in this launch: only "the copy has left shared memory" matters per step
                bulk_store(p.ws + lay.step(b, k) + kSaveX + kSaveSS, sG, (uint32_t)kSaveG);
                bulk_commit();
                bulk_wait_read();
                tc::mbar_arrive(&bar_st);
            } __syncwarp(); }
        }
        if (is_st) { if (tc::elect_one()) bulk_wait_all(); __syncwarp(); }
        __syncthreads();
        if (is_epi && half == 0) s_red[ch][r] = loss_acc;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.0f;
            for (int i = 0; i < 64; ++i) tot += s_red[0][i] + s_red[1][i];
            deposit_loss(tot, b);
        }
    }
    if (rprof && (role < 2 || role >= P)) {
        const int pr = role < P ? (role < 2 ? role : 1) : (role == P ? 2 : 3);
        g_wf_prof[pr * 4 + 0] += (unsigned long long)(clock64() - c_begin);
        g_wf_prof[pr * 4 + 1] += (unsigned long long)c_ld;
        g_wf_prof[pr * 4 + 2] += (unsigned long long)c_mma;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------
// backward, one CTA per batch element, steps k = T-2 .. 1 (t = k - 1)
//   state   Y_k = [dL_k ; dR_k^T], carried WITHOUT the factor dloss / (B N N) (everything up to the softmax backward is linear in it)
//   (a)     [dS'_{k-1} | .] = L_{k-1}^T dL_k      (lanes 0-63),      [. | dS_{k-1}] = dR_k R_{k-1}^T   (lanes 64-127)
//   (b)     T2 = coef S'_{k-1} o (dS' - rowsum) (rows of dA_{k-1}^T),  T1 = coef S_{k-1} o (dS - rowsum) + dA_in / tau (rows of dA_{k-1})
//   (c)     dE_{k-1} += T1 E_k + T2^T E_k,   dE_k += T2 E_{k-1} + T1^T E_{k-1}     (frame j lives in lane half j & 1 of accumulator j & 1)
//   (d)     Y_{k-1} = [G_{k-1} R_{k-1}^T ; G_{k-1}^T L_{k-1}] + Y_k [S'_{k-1}^T | S_{k-1}]
//   (e)     dx_k = (dE_k - E_k (E_k . dE_k)) / |x_k|     (F.normalize backward), frame k is complete after step k
// Every transposed operand is the SAME tile read through the other UMMA majorness; an MN-major A whose rows must land in lanes
// 64-127 is addressed one 64-element group (8 KB) below its tile (the group that falls on lanes 0-63 multiplies junk into
// accumulator rows nobody reads).
// ------------------------------------------------------------------------------------------
constexpr uint32_t bX = 0;                               // X_{k-1}: hi, lo                      } one 80 KB block of the saved
constexpr uint32_t bSS = bX + 2 * kPlane128;             // rows of S_{k-1}: hi, lo; S'_{k-1}: hi, lo } workspace, loaded by one
constexpr uint32_t bG = bSS + 4 * kTile64;               // rows of G_{k-1}: hi, lo               } bulk copy
constexpr uint32_t bY = bG + 2 * kTile64;                // Y = [dL ; dR^T]: hi, lo
constexpr uint32_t bT = bY + 2 * kPlane128;              // [T1 ; T2] (or [T2 ; T1]): hi, lo
constexpr uint32_t bE = bT + 2 * kPlane128;              // E planes [hi, lo][k-block][128 rows = slot 0, slot 1][128 B]
constexpr uint32_t bEnd = bE + 4 * kPlane128;
constexpr uint32_t kSmemBwd = bEnd + 1024;
static_assert(kSaveX == 2 * kPlane128 && kSaveSS == 4 * kTile64 && kSaveG == 2 * kTile64, "the saved block is the smem map bX .. bG");

struct BwdParams {
    int B, T, N, prof;
    float inv_tau;
    const float* x;           // [B, T, N, 128] raw encoder output
    const float* dloss;       // scalar
    const float* dA;          // [B, T-1, N, N] gradient flowing into the returned A, or null
    float* dx;                // [B, T, N, 128]
    const uint8_t* ws;        // saved workspace of the forward kernel (1024-aligned)
};

__global__ void __launch_bounds__(kBwdThreads, 1) walk_fused_bwd_kernel(BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.x, N = p.N, T = p.T, K = T - 2;
    const Layout lay(p.B, T);
    const float* invn = reinterpret_cast<const float*>(p.ws + (size_t)b * lay.per_b + lay.invn);

    if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_ld, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const bool is_epi = warp < 4, is_iss = warp == 4;
    const uint32_t tmem = tmem_base_s;
    const uint32_t tDS = tmem, tY = tmem + 128u, tE0 = tmem + 256u;      // dE accumulators: tE0 + 128 (j & 1)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ld_phase = 0, mma_phase = 0;
    const int half = tid >> 6, r = tid & 63;

    const uint32_t X_hi = sb + bX, X_lo = X_hi + kPlane128;
    const uint32_t sST_hi = sb + bSS, sST_lo = sST_hi + kTile64, sSp_hi = sST_hi + 2 * kTile64, sSp_lo = sST_hi + 3 * kTile64;
    const uint32_t G_hi = sb + bG, G_lo = G_hi + kTile64;
    const uint32_t Y_hi = sb + bY, Y_lo = Y_hi + kPlane128;
    const uint32_t T_hi = sb + bT, T_lo = T_hi + kPlane128;
    auto ePlane = [&](int pl, int kb) { return sb + bE + (uint32_t)(pl * 2 + kb) * kPlane128; };
    const float scale = __ldg(p.dloss) / ((float)p.B * (float)N * (float)N);
    const float coef = scale * p.inv_tau;

    auto load_block = [&](int j) {           // thread 0: [X_j | S_j, S'_j | G_j] -> bX .. bG
        tc::mbar_arrive_expect_tx(&bar_ld, (uint32_t)kSaveStep);
        bulk_load(sb + bX, p.ws + lay.step(b, j), (uint32_t)kSaveStep, &bar_ld);
    };
    // frame f: x * (1 / |x|) -> bf16 hi / lo rows of ring slot f & 1; thread = (channel half h, row r)
    auto convert_frame = [&](int f) {
        if (!is_epi) return;
        float v[64];
        if (r < N) {
            const float inv = invn[f * 64 + r];
            const float4* src = reinterpret_cast<const float4*>(p.x + ((size_t)(b * T + f) * N + r) * 128 + 64 * half);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 q = __ldg(src + i);
                v[4 * i] = q.x * inv; v[4 * i + 1] = q.y * inv; v[4 * i + 2] = q.z * inv; v[4 * i + 3] = q.w * inv;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = 0.0f;
        }
        store_row64(ePlane(0, half), ePlane(1, half), (f & 1) * 64 + r, v);
    };
    // dE_j (complete, in lane half j & 1 of accumulator j & 1) -> dx_j: F.normalize backward, row r of frame j
    auto finish_frame = [&](int j) {
        if (!is_epi || half != (j & 1)) return;                          // (warp-uniform)
        float g[128];
        const uint32_t ta = tE0 + (uint32_t)((j & 1) * 128) + lane_base;
#pragma unroll
        for (int c = 0; c < 4; ++c) tc::tmem_ld_32x32b_x32(ta + (uint32_t)(c * 32), reinterpret_cast<float(&)[32]>(g[c * 32]));
        tc::tmem_ld_wait();
        if (r >= N) return;
        const float inv = invn[j * 64 + r];
        const float4* xr = reinterpret_cast<const float4*>(p.x + ((size_t)(b * T + j) * N + r) * 128);
        float dot = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float4 q = __ldg(xr + i);
            dot = fmaf(q.x, g[4 * i], dot); dot = fmaf(q.y, g[4 * i + 1], dot); dot = fmaf(q.z, g[4 * i + 2], dot); dot = fmaf(q.w, g[4 * i + 3], dot);
        }
        dot *= inv * inv;                                                // E . dE = inv (x . dE); the factor E = inv x below takes the other inv
        float4* dst = reinterpret_cast<float4*>(p.dx + ((size_t)(b * T + j) * N + r) * 128);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float4 q = __ldg(xr + i);
            float4 o;
            o.x = (g[4 * i] - q.x * dot) * inv; o.y = (g[4 * i + 1] - q.y * dot) * inv;
            o.z = (g[4 * i + 2] - q.z * dot) * inv; o.w = (g[4 * i + 3] - q.w * dot) * inv;
            dst[i] = o;
        }
    };
    // (c): the two dE products of step k from the T tile (T1 rows in tile half (k-1) & 1, T2 rows in the other half)
    auto issue_dE = [&](int k, bool fresh_k) {
        const int p1 = (k - 1) & 1, p2 = k & 1;
        const uint32_t T1o = (uint32_t)p1 * kTile64, T2o = (uint32_t)p2 * kTile64;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const uint32_t Ek_hi = ePlane(0, kb) + (uint32_t)p2 * kTile64, Ek_lo = ePlane(1, kb) + (uint32_t)p2 * kTile64;
            const uint32_t Em_hi = ePlane(0, kb) + (uint32_t)p1 * kTile64, Em_lo = ePlane(1, kb) + (uint32_t)p1 * kTile64;
            const uint32_t acc1 = tE0 + (uint32_t)(p1 * 128 + kb * 64), acc2 = tE0 + (uint32_t)(p2 * 128 + kb * 64);
            mma3<false, true>(acc1, T_hi, T_lo, Ek_hi, Ek_lo, 1, 0, 0, 0, true);                                                   // T1 E_k
            mma3<true, true>(acc1, T_hi + T2o - (uint32_t)p1 * kTile64, T_lo + T2o - (uint32_t)p1 * kTile64, Ek_hi, Ek_lo, 1, 0, 0, kTile64, false);   // T2^T E_k
            mma3<false, true>(acc2, T_hi, T_lo, Em_hi, Em_lo, 1, 0, 0, 0, fresh_k);                                                // T2 E_{k-1}
            mma3<true, true>(acc2, T_hi + T1o - (uint32_t)p2 * kTile64, T_lo + T1o - (uint32_t)p2 * kTile64, Em_hi, Em_lo, 1, 0, 0, kTile64, false);   // T1^T E_{k-1}
        }
    };
    auto mma_wait = [&]() {
        tc::mbar_wait(&bar_mma, mma_phase & 1);
        ++mma_phase;
        tc::tc_fence_after();
    };
    auto publish = [&]() {                   // generic-proxy writes of operand tiles / TMEM reads done -> visible to the next MMAs
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
    };

    const int k_first = p.dA ? K + 1 : K;
    if (is_iss && tc::elect_one()) load_block(K);
    convert_frame(k_first);
    convert_frame(k_first - 1);
    if (!p.dA) {                             // frame T-1 only enters the last affinity, which the loss does not see
        float4* dst = reinterpret_cast<float4*>(p.dx + (size_t)(b * T + T - 1) * N * 128);
        for (int i = tid; i < N * 32; i += kBwdThreads) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::mbar_wait(&bar_ld, ld_phase & 1); ++ld_phase;
    publish();
    // Y_K = [G_K R_K^T ; G_K^T L_K]
    if (is_iss && tc::elect_one()) {
        tc::tc_fence_after();
        mma3<false, true>(tY, G_hi, G_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, 0, true);
        mma3<true, true>(tY + 64u, G_hi - kTile64, G_lo - kTile64, X_hi, X_lo, 1, 0, 0, kTile64, true);
        tc::umma_commit(&bar_mma);
    }
    mma_wait();
    if (is_epi) {
        float y[64];
        tmem_ld64(tY + lane_base + (uint32_t)(half * 64), y);
        store_row64(Y_hi, Y_lo, tid, y);
    }
    publish();
    if (is_iss && tc::elect_one()) load_block(K - 1);
    if (p.dA) {
        // pseudo-step k = K + 1 (t = T - 2): only the gradient that arrives through the returned A
        float v[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] = 0.0f;
        const int p1 = K & 1, p2 = (K + 1) & 1;
        if (!is_epi) {
        } else if (half == 1) {
            if (r < N) {
                const float* src = p.dA + (((size_t)b * (T - 1) + (T - 2)) * N + r) * N;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < N) v[c] = p.inv_tau * __ldg(src + c);
            }
            store_row64(T_hi, T_lo, p1 * 64 + r, v);
        } else {
            store_row64(T_hi, T_lo, p2 * 64 + r, v);
        }
        publish();
        if (is_iss && tc::elect_one()) {
            tc::tc_fence_after();
            issue_dE(K + 1, true);
            tc::umma_commit(&bar_mma);
        }
        mma_wait();
        finish_frame(K + 1);
        tc::tc_fence_before();
        __syncthreads();
        convert_frame(K - 1);                // the ring now holds frames K and K - 1, as step K expects
    }
    tc::mbar_wait(&bar_ld, ld_phase & 1); ++ld_phase;
    publish();

    long long c_last = clock64();
    for (int k = K; k >= 1; --k) {
        const int p1 = (k - 1) & 1, p2 = k & 1;
        // ---- (a) ----
        if (k >= 2) {
            if (is_iss && tc::elect_one()) {
                tc::tc_fence_after();
                mma3<true, true>(tDS, X_hi, X_lo, Y_hi, Y_lo, 1, 0, 0, kTile64, true);                              // L_{k-1}^T dL_k
                mma3<true, true>(tDS + 64u, Y_hi, Y_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, kTile64, true);    // dR_k R_{k-1}^T
                tc::umma_commit(&bar_mma);
            }
            WFPROF(1, 0);
            mma_wait();
            WFPROF(1, 1);
        }
        // ---- (b) ----
        if (is_epi) {
            float d[64], P[64];
            if (k >= 2) {
                tmem_ld64(tDS + lane_base + (uint32_t)(half * 64), d);
            } else if (half == 0) {
                load_row64(Y_hi, Y_lo, r, d);                             // L_0 = I: dS'_0 = dL_1
            } else {
#pragma unroll
                for (int c = 0; c < 64; ++c) d[c] = 0.0f;                 // R_1 = I is not a product: dS_0 = 0
            }
            if (half == 0) load_row64(sSp_hi, sSp_lo, r, P);
            else load_row64(sST_hi, sST_lo, r, P);
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < 64; ++c) s = fmaf(P[c], d[c], s);
#pragma unroll
            for (int c = 0; c < 64; ++c) d[c] = coef * P[c] * (d[c] - s);
            if (half == 1 && p.dA && r < N) {
                const float* src = p.dA + (((size_t)b * (T - 1) + (k - 1)) * N + r) * N;
#pragma unroll
                for (int c = 0; c < 64; ++c)
                    if (c < N) d[c] = fmaf(p.inv_tau, __ldg(src + c), d[c]);
            }
            store_row64(T_hi, T_lo, (half == 1 ? p1 : p2) * 64 + r, d);
        }
        publish();
        WFPROF(1, 2);
        // ---- (c), (d) ----
        if (is_iss && tc::elect_one()) {
            tc::tc_fence_after();
            issue_dE(k, k == k_first);
            if (k >= 2) {
                mma3<false, false>(tY, Y_hi, Y_lo, sSp_hi, sSp_lo, 1, 0, 0, 0, true);                                  // dL_k S'_{k-1}^T
                mma3<false, true>(tY, G_hi, G_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, 0, false);                 // G_{k-1} R_{k-1}^T
                mma3<false, true>(tY + 64u, Y_hi, Y_lo, sST_hi, sST_lo, 1, 0, 0, 0, true);                             // dR_k^T S_{k-1}
                mma3<true, true>(tY + 64u, G_hi - kTile64, G_lo - kTile64, X_hi, X_lo, 1, 0, 0, kTile64, false);      // G_{k-1}^T L_{k-1}
            }
            tc::umma_commit(&bar_mma);
        }
        WFPROF(1, 3);
        mma_wait();
        WFPROF(1, 4);
        // ---- (e) ----
        if (k >= 2 && is_epi) {
            float y[64];
            tmem_ld64(tY + lane_base + (uint32_t)(half * 64), y);
            store_row64(Y_hi, Y_lo, tid, y);
        }
        finish_frame(k);
        if (k == 1) finish_frame(0);
        publish();
        WFPROF(1, 5);
        // ---- next step's tiles and frame ----
        if (k >= 2) {
            if (is_iss && tc::elect_one()) load_block(k - 2);
            convert_frame(k - 2);
            WFPROF(1, 6);
            tc::mbar_wait(&bar_ld, ld_phase & 1); ++ld_phase;
            publish();
            WFPROF(1, 7);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// backward, role-split: 2 + PE CTAs per batch element (grid = (2 + PE) B, role = blockIdx.x / B), hand-over through L2 as in the
// forward.  Same algebra as walk_fused_bwd_kernel:
//   role 0        chain: Y_K = own_K, Y_{k-1} = own_{k-1} + Y_k [S'_{k-1}^T | S_{k-1}]; saves every Y_k + flag (one product pair and one
//                 32-column epilogue per step: the only serial part)
//   role 1        dA: per step (a) [dS' | dS] from Y_k and X_{k-1}, (b) softmax backward -> the tile [T1 ; T2] (rows of dA_{k-1} and of
//                 its transpose, scaled by dloss / (B N N tau), plus the gradient arriving through the returned A) + flag
//   roles 2 ..    dE: per FRAME j: dE_j = T1_{j+1} E_{j+1} + T2_{j+1}^T E_{j+1} + T2_j E_{j-1} + T1_j^T E_{j-1} (four N = 128 products on the
//                 frames' operand rows the forward saved), F.normalize backward, dx_j.  Frames are dealt round-robin, last frame first.
// ------------------------------------------------------------------------------------------
struct BScratch {            // backward scratch (bytes): per batch element Y_k (k = 1 .. K) and T_k (k = 1 .. K + 1) tiles, then the flags
    size_t per_b, flags, total;
    int T, K;
    __host__ __device__ BScratch(int B, int T_) : T(T_), K(T_ - 2) {
        per_b = (size_t)(2 * K + 1) * kSaveX;
        flags = (size_t)B * per_b;
        total = flags + align_up((size_t)(B * 2 * T_ + 1) * sizeof(int), 256);
    }
    __host__ __device__ size_t Y(int b, int k) const { return (size_t)b * per_b + (size_t)(k - 1) * kSaveX; }
    __host__ __device__ size_t Tt(int b, int k) const { return (size_t)b * per_b + (size_t)(K + k - 1) * kSaveX; }
    __host__ __device__ size_t flag(int b, int which, int i) const { return flags + ((size_t)(b * 2 + which) * T + i + 1) * sizeof(int); }   // [0] = error flag
};

struct BwdRolesParams {
    BwdParams q;
    uint8_t* scratch;         // BScratch (1024-aligned)
    int PE;                   // dE CTAs per batch element
    int dbg;                  // development switches (CRW_WALK_DBG)
};

constexpr uint32_t kSmemBwdRoles = 7 * 2 * kPlane128 + 1024;             // 224 KB (the dA role's map) + alignment slack

__global__ void __launch_bounds__(kRoleThreads, 1) walk_fused_bwd_roles_kernel(BwdRolesParams pp) {
    const BwdParams& p = pp.q;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_ld[2], bar_mma, bar_st, bar_tile;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_red[4][64];
    __shared__ int s_next;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = blockIdx.x / p.B, b = blockIdx.x % p.B, N = p.N, T = p.T, K = T - 2;
    const Layout lay(p.B, T);
    const BScratch sc(p.B, T);
    int* err = reinterpret_cast<int*>(pp.scratch + sc.flags);
    auto flagp = [&](int which, int i) { return reinterpret_cast<int*>(pp.scratch + sc.flag(b, which, i)); };
    const float* invn = reinterpret_cast<const float*>(p.ws + (size_t)b * lay.per_b + lay.invn);

    if (warp == 0) tc::tmem_alloc<128>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_ld[0], 1);
        tc::mbar_init(&bar_ld[1], 1);
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_init(&bar_st, 1);
        tc::mbar_init(&bar_tile, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const bool is_epi = warp < 8, is_iss = warp == 8, is_st = warp == 9;      // (warp 9: bulk stores, their completion, the flags)
    const uint32_t tmem = tmem_base_s;
    const int row = ((warp & 3) << 5) | lane;
    const int half = row >> 6, r = row & 63, ch = (warp >> 2) & 1;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ld_phase[2] = {0u, 0u}, mma_phase = 0;
    // the store warp's copy of a tile has LEFT shared memory (arrives on bar_st): waited for by everybody right before the tile is
    // overwritten in place; its completion in global memory and the flag follow on the store lane alone
    uint32_t st_phase = 0;
    auto st_wait = [&]() {
        if (is_st) return;                   // (the store warp is the one that arrives; see ld_wait)
        tc::mbar_wait(&bar_st, st_phase & 1);
        ++st_phase;
    };
    auto store_tile = [&](void* gdst, uint32_t ssrc, uint32_t bytes, int* flag) {     // store lane
        bulk_store(gdst, ssrc, bytes);
        bulk_commit();
        bulk_wait_read();
        tc::mbar_arrive(&bar_st);
        bulk_wait_all();
        if (flag) flag_set(flag);
    };
    auto publish = [&]() {
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
    };
    // profiling aid (CRW_WALK_PROF=1): per role of element 0, cycles [total, waiting for loads / hand-over, waiting for MMAs]
    const bool rprof = p.prof && b == 0 && tid == 0;
    const long long c_begin = clock64();
    long long c_ld = 0, c_mma = 0;
    auto mma_wait = [&]() {
        if (is_st) return;
        const long long c0 = rprof ? clock64() : 0;
        tc::mbar_wait(&bar_mma, mma_phase & 1);
        ++mma_phase;
        tc::tc_fence_after();
        if (rprof) c_mma += clock64() - c0;
    };
    // (The store warp skips the waits on load / MMA barriers: it needs none of that data, and, coming late from a copy it completed,
    // it could find a barrier re-armed and completed AGAIN -- same parity -- and wait for ever.)
    auto ld_wait = [&](int i) {
        if (is_st) return;
        const long long c0 = rprof ? clock64() : 0;
        tc::mbar_wait(&bar_ld[i], ld_phase[i] & 1);
        ++ld_phase[i];
        if (rprof) c_ld += clock64() - c0;
    };
    const int k_first = p.dA ? K + 1 : K;

    if (((pp.dbg & 2) && role >= 1) || ((pp.dbg & 4) && role >= 2)) {
        // (development: role switched off)
    } else if (role == 0) {
        // ================= chain =================
        // (the serial role; the store warp runs on its own as in the forward's chain: bar_tile "Y is written" -> copy -> bar_st "the copy
        // has left shared memory" -> completion -> flag)
        const uint32_t sY = sb + 2 * (uint32_t)kSaveStep;                // blocks: two 80 KB buffers at 0; Y: hi, lo
        const uint32_t Y_hi = sY, Y_lo = sY + kPlane128;
        if (is_st) {
            for (int k = K; k >= 1; --k) {
                tc::mbar_wait(&bar_tile, (K - k) & 1);
                if (tc::elect_one()) store_tile(pp.scratch + sc.Y(b, k), sY, (uint32_t)kSaveX, flagp(0, k));
                __syncwarp();
            }
        } else {
            auto sync9 = [&]() {
                tc::fence_proxy_async();
                tc::tc_fence_before();
                asm volatile("bar.sync 1, 288;" ::: "memory");
            };
            auto blk = [&](int j) { return sb + (uint32_t)(j & 1) * (uint32_t)kSaveStep; };
            auto load_block = [&](int j) {               // issuer lane: [X_j | S_j, S'_j | G_j] of the forward's workspace
                tc::mbar_arrive_expect_tx(&bar_ld[j & 1], (uint32_t)kSaveStep);
                bulk_load(blk(j), p.ws + lay.step(b, j), (uint32_t)kSaveStep, &bar_ld[j & 1]);
            };
            if (is_iss) { if (tc::elect_one()) { load_block(K); if (K >= 2) load_block(K - 1); } __syncwarp(); }
            ld_wait(K & 1);
            if (is_iss) { if (tc::elect_one()) {             // Y_K = [G_K R_K^T ; G_K^T L_K]
                const uint32_t X_hi = blk(K), X_lo = X_hi + kPlane128, G_hi = X_hi + (uint32_t)(kSaveX + kSaveSS), G_lo = G_hi + kTile64;
                tc::tc_fence_after();
                mma3<false, true>(tmem, G_hi, G_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, 0, true);
                mma3<true, true>(tmem + 64u, G_hi - kTile64, G_lo - kTile64, X_hi, X_lo, 1, 0, 0, kTile64, true);
                tc::umma_commit(&bar_mma);
            } __syncwarp(); }
            mma_wait();
            if (is_epi) {
                float y[32];
                tmem_ld32(tmem + lane_base + (uint32_t)(half * 64 + ch * 32), y);
                store_row32(Y_hi, Y_lo, row, ch, y);
            }
            sync9();
            if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
            for (int k = K; k >= 2; --k) {
                ld_wait((k - 1) & 1);
                if (is_iss) { if (tc::elect_one()) {         // Y_{k-1} = own_{k-1} + Y_k [S'_{k-1}^T | S_{k-1}]
                    const uint32_t X_hi = blk(k - 1), X_lo = X_hi + kPlane128, S_hi = X_hi + (uint32_t)kSaveX, G_hi = S_hi + (uint32_t)kSaveSS, G_lo = G_hi + kTile64;
                    tc::tc_fence_after();
                    mma3<false, false>(tmem, Y_hi, Y_lo, S_hi + 2 * kTile64, S_hi + 3 * kTile64, 1, 0, 0, 0, true);             // dL_k S'_{k-1}^T
                    mma3<false, true>(tmem, G_hi, G_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, 0, false);                    // G_{k-1} R_{k-1}^T
                    mma3<false, true>(tmem + 64u, Y_hi, Y_lo, S_hi, S_hi + kTile64, 1, 0, 0, 0, true);                          // dR_k^T S_{k-1}
                    mma3<true, true>(tmem + 64u, G_hi - kTile64, G_lo - kTile64, X_hi, X_lo, 1, 0, 0, kTile64, false);         // G_{k-1}^T L_{k-1}
                    tc::umma_commit(&bar_mma);
                    if (k >= 3) load_block(k - 2);       // (its buffer was read by the products of the step before: complete)
                } __syncwarp(); }
                mma_wait();
                st_wait();                                                   // the copy of Y_k has left shared memory
                if (is_epi) {
                    float y[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)(half * 64 + ch * 32), y);
                    store_row32(Y_hi, Y_lo, row, ch, y);
                }
                sync9();
                if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
            }
        }
    } else if (role == 1) {
        // ================= dA =================
        const uint32_t sYb = sb, sXS = sb + 2 * (uint32_t)kSaveX, sT = sXS + 2 * (uint32_t)(kSaveX + kSaveSS);   // Y x 2 | (X, S) x 2 | T
        const uint32_t T_hi = sT, T_lo = sT + kPlane128;
        const float scale = __ldg(p.dloss) / ((float)p.B * (float)N * (float)N);
        const float coef = scale * p.inv_tau;
        auto load_y = [&](int k) {                   // issuer lane; Y_k -> buffer k & 1 (barrier 0)
            flag_wait(flagp(0, k), err);
            tc::mbar_arrive_expect_tx(&bar_ld[0], (uint32_t)kSaveX);
            bulk_load(sYb + (uint32_t)(k & 1) * (uint32_t)kSaveX, pp.scratch + sc.Y(b, k), (uint32_t)kSaveX, &bar_ld[0]);
        };
        auto load_xs = [&](int j) {                  // issuer lane; [X_j | S_j, S'_j] -> buffer j & 1 (barrier 1)
            tc::mbar_arrive_expect_tx(&bar_ld[1], (uint32_t)(kSaveX + kSaveSS));
            bulk_load(sXS + (uint32_t)(j & 1) * (uint32_t)(kSaveX + kSaveSS), p.ws + lay.step(b, j), (uint32_t)(kSaveX + kSaveSS), &bar_ld[1]);
        };
        if (is_st) {
            // the store warp on its own: "T_k is written" (bar_tile) -> copy -> "the copy has left shared memory" (bar_st) -> completion -> flag
            int n = 0;
            for (int k = k_first; k >= 1; --k, ++n) {
                tc::mbar_wait(&bar_tile, n & 1);
                if (tc::elect_one()) store_tile(pp.scratch + sc.Tt(b, k), sT, (uint32_t)kSaveX, flagp(1, k));       // the dE CTAs are waiting for T_k
                __syncwarp();
            }
        } else {
            auto sync9 = [&]() {
                tc::fence_proxy_async();
                tc::tc_fence_before();
                asm volatile("bar.sync 1, 288;" ::: "memory");
            };
            int y_low = K + 1;                       // (issuer lane) Y_k has been fetched for k >= y_low
            if (is_iss) { if (tc::elect_one()) load_xs(K - 1); __syncwarp(); }
            if (p.dA) {
                // pseudo-step k = K + 1 (t = T - 2): only the gradient that arrives through the returned A
                if (is_epi) {
                    float v[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = 0.0f;
                    if (half == 1 && r < N) {
                        const float* src = p.dA + (((size_t)b * (T - 1) + (T - 2)) * N + r) * N;
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (32 * ch + c < N) v[c] = p.inv_tau * __ldg(src + 32 * ch + c);
                    }
                    store_row32(T_hi, T_lo, (half == 1 ? 0 : 64) + r, ch, v);
                }
                sync9();
                if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
            }
            for (int k = K; k >= 1; --k) {
                const uint32_t Y_hi = sYb + (uint32_t)(k & 1) * (uint32_t)kSaveX, Y_lo = Y_hi + kPlane128;
                const uint32_t X_hi = sXS + (uint32_t)((k - 1) & 1) * (uint32_t)(kSaveX + kSaveSS), X_lo = X_hi + kPlane128, S_hi = X_hi + (uint32_t)kSaveX;
                if (is_iss) { if (tc::elect_one()) { if (y_low > k) { load_y(k); y_low = k; } } __syncwarp(); }       // (waits for the flag here)
                ld_wait(0);
                ld_wait(1);
                if (is_iss) { if (tc::elect_one()) {
                    tc::tc_fence_after();
                    if (k >= 2) {
                        mma3<true, true>(tmem, X_hi, X_lo, Y_hi, Y_lo, 1, 0, 0, kTile64, true);                              // L_{k-1}^T dL_k
                        mma3<true, true>(tmem + 64u, Y_hi, Y_lo, X_hi + kTile64, X_lo + kTile64, 1, 0, 0, kTile64, true);    // dR_k R_{k-1}^T
                    }
                    tc::umma_commit(&bar_mma);
                    if (k >= 2) {
                        load_xs(k - 2);
                        if (flag_peek(flagp(0, k - 1))) { load_y(k - 1); y_low = k - 1; }    // (only if it is already there)
                    }
                } __syncwarp(); }
                mma_wait();
                float d[32], P[32];
                float s = 0.0f;
                if (is_epi) {
                    if (k >= 2) {
                        tmem_ld32(tmem + lane_base + (uint32_t)(half * 64 + ch * 32), d);
                    } else if (half == 0) {
                        load_row32(Y_hi, Y_lo, r, ch, d);                     // L_0 = I: dS'_0 = dL_1
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) d[c] = 0.0f;             // R_1 = I is not a product: dS_0 = 0
                    }
                    if (half == 0) load_row32(S_hi + 2 * kTile64, S_hi + 3 * kTile64, r, ch, P);
                    else load_row32(S_hi, S_hi + kTile64, r, ch, P);
#pragma unroll
                    for (int c = 0; c < 32; ++c) s = fmaf(P[c], d[c], s);
                    s_red[half * 2 + ch][r] = s;
                }
                tc::tc_fence_before();
                asm volatile("bar.sync 1, 288;" ::: "memory");
                if (k < K || p.dA) st_wait();                                // the copy of T_{k+1} has left shared memory
                if (is_epi) {
                    s += s_red[half * 2 + (1 - ch)][r];
#pragma unroll
                    for (int c = 0; c < 32; ++c) d[c] = coef * P[c] * (d[c] - s);
                    if (half == 1 && p.dA && r < N) {
                        const float* src = p.dA + (((size_t)b * (T - 1) + (k - 1)) * N + r) * N;
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (32 * ch + c < N) d[c] = fmaf(p.inv_tau, __ldg(src + 32 * ch + c), d[c]);
                    }
                    store_row32(T_hi, T_lo, (half == 1 ? 0 : 64) + r, ch, d);   // T1 (rows of dA) on top, T2 (rows of dA^T) below
                }
                sync9();
                if (is_iss) { if (tc::elect_one()) tc::mbar_arrive(&bar_tile); __syncwarp(); }
            }
        }
    } else {
        // ================= dE (frames) =================
        const int e = role - 2;
        const uint32_t sTa = sb, sTb = sb + (uint32_t)kSaveX, sEa = sb + 2 * (uint32_t)kSaveX, sEb = sb + 3 * (uint32_t)kSaveX;
        if (e == 0 && !p.dA) {                       // frame T-1 only enters the last affinity, which the loss does not see
            float4* dst = reinterpret_cast<float4*>(p.dx + (size_t)(b * T + T - 1) * N * 128);
            for (int i = tid; i < N * 32; i += kRoleThreads) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        auto load_frame_tiles = [&](int j) {         // issuer lane: T_{j+1}, E_{j+1}, T_j, E_{j-1} (those that exist)
            const bool has1 = j + 1 <= k_first, has2 = j >= 1;
            const uint32_t bytes = (uint32_t)((has1 ? 1 : 0) + (has2 ? 1 : 0)) * (uint32_t)(kSaveX + kSaveFrame);
            if (has1) flag_wait(flagp(1, j + 1), err);
            if (has2) flag_wait(flagp(1, j), err);
            tc::mbar_arrive_expect_tx(&bar_ld[0], bytes);
            if (has1) {
                bulk_load(sTa, pp.scratch + sc.Tt(b, j + 1), (uint32_t)kSaveX, &bar_ld[0]);
                bulk_load(sEa, p.ws + lay.frame(b, j + 1), (uint32_t)kSaveFrame, &bar_ld[0]);
            }
            if (has2) {
                bulk_load(sTb, pp.scratch + sc.Tt(b, j), (uint32_t)kSaveX, &bar_ld[0]);
                bulk_load(sEb, p.ws + lay.frame(b, j - 1), (uint32_t)kSaveFrame, &bar_ld[0]);
            }
        };
        bool have_next = false;                      // this frame's tiles were fetched under the previous frame's epilogue
        for (int j = k_first - e; j >= 0; j -= pp.PE) {
            const bool has1 = j + 1 <= k_first, has2 = j >= 1;
            if (!have_next && is_iss) { if (tc::elect_one()) load_frame_tiles(j); __syncwarp(); }
            ld_wait(0);
            if (is_iss) { if (tc::elect_one()) {
                tc::tc_fence_after();
                // frame tiles: [hi, lo][k-block][64 rows][128 B] -> MN-major B with two 64-channel groups 8 KB apart
                if (has1) {
                    mma3<false, true, 128>(tmem, sTa, sTa + kPlane128, sEa, sEa + 2 * kTile64, 1, 0, 0, 0, true, kTile64);                       // T1 E_{j+1}
                    mma3<true, true, 128>(tmem, sTa + kTile64, sTa + kPlane128 + kTile64, sEa, sEa + 2 * kTile64, 1, 0, 0, kTile64, false, kTile64); // T2^T E_{j+1}
                }
                if (has2) {
                    mma3<false, true, 128>(tmem, sTb + kTile64, sTb + kPlane128 + kTile64, sEb, sEb + 2 * kTile64, 1, 0, 0, 0, !has1, kTile64);  // T2 E_{j-1}
                    mma3<true, true, 128>(tmem, sTb, sTb + kPlane128, sEb, sEb + 2 * kTile64, 1, 0, 0, kTile64, false, kTile64);                  // T1^T E_{j-1}
                }
                tc::umma_commit(&bar_mma);
            } __syncwarp(); }
            mma_wait();
            // the operand buffers are free: fetch the next frame's tiles under this frame's epilogue
            // (only when they are already there: the issuer must not sit in a flag wait while this frame's epilogue needs it at a barrier)
            bool fetched = false;
            if (j - pp.PE >= 0) {
                const int jn = j - pp.PE;
                int ok = 0;
                if (is_iss) {
                    if (tc::elect_one()) {
                        ok = (jn + 1 > k_first || flag_peek(flagp(1, jn + 1))) && (jn < 1 || flag_peek(flagp(1, jn)));
                        if (ok) load_frame_tiles(jn);
                    }
                    ok = __shfl_sync(0xffffffffu, ok, 0) | __any_sync(0xffffffffu, ok);
                    if (lane == 0) s_next = ok;
                }
                fetched = true;
            }
            // F.normalize backward on the rows in lanes 0-63; thread = (row r, 64 channels)
            float g[64];
            float dot = 0.0f;
            const bool mine = is_epi && half == 0;
            const float inv = (mine && r < N) ? invn[j * 64 + r] : 0.0f;
            const float4* xr = reinterpret_cast<const float4*>(p.x + ((size_t)(b * T + j) * N + (r < N ? r : 0)) * 128 + 64 * ch);
            if (mine) {
                tmem_ld64(tmem + lane_base + (uint32_t)(ch * 64), g);
                if (r < N) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 q = __ldg(xr + i);
                        dot = fmaf(q.x, g[4 * i], dot); dot = fmaf(q.y, g[4 * i + 1], dot); dot = fmaf(q.z, g[4 * i + 2], dot); dot = fmaf(q.w, g[4 * i + 3], dot);
                    }
                }
                s_red[ch][r] = dot;
            }
            tc::tc_fence_before();
            __syncthreads();
            if (mine && r < N) {
                dot = (dot + s_red[1 - ch][r]) * inv * inv;
                float4* dst = reinterpret_cast<float4*>(p.dx + ((size_t)(b * T + j) * N + r) * 128 + 64 * ch);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 q = __ldg(xr + i);
                    float4 o;
                    o.x = (g[4 * i] - q.x * dot) * inv; o.y = (g[4 * i + 1] - q.y * dot) * inv;
                    o.z = (g[4 * i + 2] - q.z * dot) * inv; o.w = (g[4 * i + 3] - q.w * dot) * inv;
                    dst[i] = o;
                }
            }
            tc::tc_fence_before();
            __syncthreads();
            have_next = fetched && s_next != 0;
        }
        // a hand-over flag that never arrived (this launch or the forward's) must not pass for a gradient: poison the element
        if (e == 0 && tid == 0 && *reinterpret_cast<volatile int*>(err)) p.dx[(size_t)b * T * N * 128] = __int_as_float(0x7fc00000);
    }
    if (rprof && role < 4) {
        g_wf_prof[16 + role * 4 + 0] += (unsigned long long)(clock64() - c_begin);
        g_wf_prof[16 + role * 4 + 1] += (unsigned long long)c_ld;
        g_wf_prof[16 + role * 4 + 2] += (unsigned long long)c_mma;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace wf

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW wf_encode_fn() {
    static EncodeTiledFnW fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnW>(p);
    }
    return fn;
}
// fp32 [rows][128] row-major, box = [box_rows][32 floats = 128 B], SWIZZLE_128B
static int make_tmap_f32_c32(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFnW fn = wf_encode_fn();
    if (!fn) return CRW_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || box_rows < 1 || box_rows > 256) return CRW_ERR_ALIGN;
    cuuint64_t gdim[2] = {128, rows};
    cuuint64_t gstr[1] = {512};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRW_OK : CRW_ERR_INVALID;
}

bool walk_fused_supported(int N, int C, int T) { return N >= 8 && N <= 64 && C == 128 && T >= 3; }
size_t walk_fused_saved_bytes(int B, int T) { return wf::Layout(B, T).total + 1024; }

// Role-split kernels (several CTAs per batch element, hand-over through L2): number of affinity producers per element, 0 = one CTA
// per element.  Used while every CTA of the launch can be resident at once (four per element); a CTA only ever waits for CTAs with
// a smaller block index and producers wait for nobody, so larger launches would not deadlock either, but they gain nothing (measured:
// four producers per element, 192 CTAs at B = 32, 49.7 vs 42 us).  CRW_WALK_ROLES=<P> forces (0 = off).
static int walk_fused_roles(int B, int sms) {
    const char* e = getenv("CRW_WALK_ROLES");
    if (e) { const int v = atoi(e); return v < 0 ? 0 : (v > 8 ? 8 : v); }
    return 4 * B <= sms ? 2 : 0;
}
bool walk_fused_roles_apply(int B) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return walk_fused_roles(B, sms) > 0;
}

static uint8_t* wf_align1k(void* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023)); }

int walk_fused_forward(const float* x, int B, int T, int N, int C, float tau, float* loss, float* A_or_null, void* saved, cudaStream_t st) {
    if (!walk_fused_supported(N, C, T)) return CRW_ERR_UNSUPPORTED;
    static bool attr_done[64] = {}, attr_done_roles[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        CRW_CUDA_RET(cudaFuncSetAttribute(wf::walk_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf::kSmemFwd));
        attr_done[dev] = true;
    }
    CUtensorMap xmap;
    int rc = make_tmap_f32_c32(&xmap, x, (uint64_t)B * T * N, (uint32_t)N);
    if (rc != CRW_OK) return rc;
    wf::FwdParams p;
    p.B = B; p.T = T; p.N = N;
    p.prof = getenv("CRW_WALK_PROF") ? 1 : 0;
    p.inv_tau = 1.0f / tau;
    p.A = A_or_null;
    p.loss = loss;
    p.ws = wf_align1k(saved);
    const wf::Layout lay(B, T);
    CRW_CUDA_RET(cudaMemsetAsync(p.ws + lay.ctr, 0, lay.total - lay.ctr, st));          // loss counter, error flag, hand-over flags
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.P = walk_fused_roles(B, sms);
    if (p.P > 0) {
        if (dev >= 0 && dev < 64 && !attr_done_roles[dev]) {
            CRW_CUDA_RET(cudaFuncSetAttribute(wf::walk_fused_fwd_roles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf::kSmemRoles));
            attr_done_roles[dev] = true;
        }
        wf::walk_fused_fwd_roles_kernel<<<(p.P + 2) * B, wf::kRoleThreads, wf::kSmemRoles, st>>>(xmap, p);
    } else {
        wf::walk_fused_fwd_kernel<<<B, wf::kThreads, wf::kSmemFwd, st>>>(xmap, p);
    }
    CRW_LAUNCH_RET();
    return CRW_OK;
}

size_t walk_fused_bwd_scratch_bytes(int B, int T) { return wf::BScratch(B, T).total + 1024; }

int walk_fused_backward(const float* x, const void* saved, const float* dloss, const float* dA_or_null, int B, int T, int N, int C, float tau,
                        float* dx, void* scratch, cudaStream_t st) {
    if (!walk_fused_supported(N, C, T)) return CRW_ERR_UNSUPPORTED;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        CRW_CUDA_RET(cudaFuncSetAttribute(wf::walk_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf::kSmemBwd));
        CRW_CUDA_RET(cudaFuncSetAttribute(wf::walk_fused_bwd_roles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf::kSmemBwdRoles));
        attr_done[dev] = true;
    }
    wf::BwdParams p;
    p.B = B; p.T = T; p.N = N;
    p.prof = getenv("CRW_WALK_PROF") ? 1 : 0;
    p.inv_tau = 1.0f / tau;
    p.x = x; p.dloss = dloss; p.dA = dA_or_null; p.dx = dx;
    p.ws = wf_align1k(const_cast<void*>(saved));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (walk_fused_roles(B, sms) > 0 && scratch) {
        wf::BwdRolesParams pp;
        pp.q = p;
        pp.scratch = wf_align1k(scratch);
        pp.PE = 2;
        { const char* e = getenv("CRW_WALK_DBG"); pp.dbg = e ? atoi(e) : 0; }
        const wf::BScratch sc(B, T);
        CRW_CUDA_RET(cudaMemsetAsync(pp.scratch + sc.flags, 0, sc.total - sc.flags, st));
        wf::walk_fused_bwd_roles_kernel<<<(2 + pp.PE) * B, wf::kRoleThreads, wf::kSmemBwdRoles, st>>>(pp);
    } else {
        wf::walk_fused_bwd_kernel<<<B, wf::kBwdThreads, wf::kSmemBwd, st>>>(p);
    }
    CRW_LAUNCH_RET();
    return CRW_OK;
}

int walk_fused_profile_read(unsigned long long* host_out, int reset) {
    if (host_out) CRW_CUDA_RET(cudaMemcpyFromSymbol(host_out, wf::g_wf_prof, sizeof(wf::g_wf_prof)));
    if (reset) {
        void* ptr = nullptr;
        CRW_CUDA_RET(cudaGetSymbolAddress(&ptr, wf::g_wf_prof));
        CRW_CUDA_RET(cudaMemset(ptr, 0, sizeof(wf::g_wf_prof)));
    }
    return CRW_OK;
}

}  // namespace crw

// profiling aid: per-phase cycle counters of the fused walk kernels ([2 kernels][16] uint64) to host memory (synchronising)
extern "C" int crw_debug_walk_fused_profile(unsigned long long* host_out, int reset) { return crw::walk_fused_profile_read(host_out, reset); }
