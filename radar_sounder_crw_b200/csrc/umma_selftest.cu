// umma_selftest.cu -- one-tile tcgen05 GEMM through exactly the TMA / descriptor / TMEM plumbing the
// tensor-core kernels use: out[128 x BN] = A[128 x 128] * B[BN x 128]^T (bf16 in, fp32 out).
// Exposed as crw_debug_umma_gemm so the GPU test-suite can pin the descriptor encodings against torch.
#include "tc_common.cuh"

namespace crw {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_bf16_k64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRW_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (cols * 2) % 16 || box_rows < 1 || box_rows > 256) return CRW_ERR_ALIGN;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRW_OK : CRW_ERR_INVALID;
}

// 2-D bf16 row-major matrix [rows][cols], box = [box_rows][box_cols elements] (box_cols * 2 bytes <= 128), SWIZZLE_128B
int make_tmap_bf16_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRW_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (cols * 2) % 16 || box_rows < 1 || box_rows > 256 || box_cols * 2 > 128)
        return CRW_ERR_ALIGN;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRW_OK : CRW_ERR_INVALID;
}

// a stack of `mats` bf16 row-major matrices [rows][cols valid of `pitch` elements], one after the other (matrix stride =
// rows * pitch); box = 64 x 64 elements of ONE matrix, SWIZZLE_128B.  Rows / columns past the matrix read as zero.
int make_tmap_bf16_mats(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t mats, uint64_t pitch) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRW_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch * 2) % 16 || cols > pitch || !cols || !rows || !mats) return CRW_ERR_ALIGN;
    cuuint64_t gdim[3] = {cols, rows, mats};
    cuuint64_t gstr[2] = {pitch * 2, rows * pitch * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CRW_OK : CRW_ERR_INVALID;
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int BN, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                     // [2 kblocks][128 rows][128 B]
    uint8_t* sB = smem + 2 * 128 * 128;     // [2 kblocks][BN rows][128 B]
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)((128 + BN) * 256));
        for (int kb = 0; kb < 2; ++kb) {
            tc::tma_load_2d(sA + kb * 128 * 128, &mapA, kb * 64, 0, &bar_full);
            tc::tma_load_2d(sB + kb * BN * 128, &mapB, kb * 64, 0, &bar_full);
        }
        tc::mbar_wait(&bar_full, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::umma_idesc_bf16(128, BN);
        for (int kb = 0; kb < 2; ++kb)
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = tc::umma_smem_desc_k128(tc::smem_u32(sA + kb * 128 * 128) + k * 32);
                const uint64_t bd = tc::umma_smem_desc_k128(tc::smem_u32(sB + kb * BN * 128) + k * 32);
                tc::umma_bf16_ss(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        float* o = out + (size_t)(warp * 32 + lane) * BN + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c + i < BN) o[i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

// Same GEMM with the A operand placed in TMEM by the threads (tcgen05.st, lane = row, two bf16 per 32-bit column)
// and consumed by the "TS" form of tcgen05.mma; B still arrives through TMA.  Pins the A-in-TMEM layout.
__global__ void __launch_bounds__(128, 1)
umma_ts_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __grid_constant__ CUtensorMap mapB, int BN, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kACol = 256;     // A occupies TMEM columns [256, 320), D columns [0, BN)

    if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)(BN * 256));
        for (int kb = 0; kb < 2; ++kb) tc::tma_load_2d(sB + kb * BN * 128, &mapB, kb * 64, 0, &bar_full);
    }
    // every thread parks its row of A (128 bf16 = 64 packed columns) in TMEM
    const uint4* arow = reinterpret_cast<const uint4*>(A + (size_t)tid * 128);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 v0 = arow[c8 * 2], v1 = arow[c8 * 2 + 1];
        const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        tc::tmem_st_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + kACol + c8 * 8, r);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        tc::mbar_wait(&bar_full, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::umma_idesc_bf16(128, BN);
        for (int k8 = 0; k8 < 8; ++k8) {
            const uint64_t bd = tc::umma_smem_desc_k128(tc::smem_u32(sB + (k8 >> 2) * BN * 128) + (k8 & 3) * 32);
            tc::umma_bf16_ts(tmem_base, tmem_base + kACol + k8 * 8, bd, idesc, k8 ? 1u : 0u);
        }
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        float* o = out + (size_t)(warp * 32 + lane) * BN + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c + i < BN) o[i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

// Same GEMM with A brought in by TMA (SWIZZLE_128B K-major tile, as the SS form reads it) and then copied into TMEM by
// tcgen05.cp (128x256b per K = 16 step), consumed by TS MMAs.  Pins the smem -> TMEM copy against the A-in-TMEM layout.
__global__ void __launch_bounds__(128, 1)
umma_tscp_selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int BN, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));   // [kb 0,1][128][128 B]
    uint8_t* sB = sA + 2 * 128 * 128;
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kACol = 256;

    if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)(2 * 128 * 128 + BN * 256));
        for (int kb = 0; kb < 2; ++kb) tc::tma_load_2d(sA + kb * 128 * 128, &mapA, kb * 64, 0, &bar_full);
        for (int kb = 0; kb < 2; ++kb) tc::tma_load_2d(sB + kb * BN * 128, &mapB, kb * 64, 0, &bar_full);
        tc::mbar_wait(&bar_full, 0);
        tc::tc_fence_after();
        for (int k8 = 0; k8 < 8; ++k8)
            tc::tmem_cp_128x256b(tmem_base + kACol + k8 * 8, tc::umma_smem_desc_k128(tc::smem_u32(sA + (k8 >> 2) * 128 * 128) + (k8 & 3) * 32));
        const uint32_t idesc = tc::umma_idesc_bf16(128, BN);
        for (int k8 = 0; k8 < 8; ++k8) {
            const uint64_t bd = tc::umma_smem_desc_k128(tc::smem_u32(sB + (k8 >> 2) * BN * 128) + (k8 & 3) * 32);
            tc::umma_bf16_ts(tmem_base, tmem_base + kACol + k8 * 8, bd, idesc, k8 ? 1u : 0u);
        }
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        float* o = out + (size_t)(warp * 32 + lane) * BN + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c + i < BN) o[i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

// MN-major operands: out[128 x BN] = A * B with A given as At[K=64][128] (M contiguous) when a_mn, else A[128][64];
// and B given as Bkn[64][BN] (N contiguous) when b_mn, else Bt[BN][64].  Pins the MN-major descriptors (LBO / SBO /
// major bits) so that transposed operands need no second copy.
int make_tmap_bf16_box(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

__global__ void __launch_bounds__(128, 1)
umma_mn_selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int BN, int a_mn, int b_mn,
                        float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;               // K-major: [128 rows][128 B] = 16 KB ; MN-major: 2 groups x [64 k][128 B] = 16 KB
    uint8_t* sB = smem + 16384;       // K-major: [BN rows][128 B]          ; MN-major: BN/64 groups x [64 k][128 B]
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)((128 + BN) * 128));
        if (a_mn) { for (int g = 0; g < 2; ++g) tc::tma_load_2d(sA + g * 8192, &mapA, g * 64, 0, &bar_full); }   // box: 64 m x 64 k
        else tc::tma_load_2d(sA, &mapA, 0, 0, &bar_full);                                                        // box: 64 k x 128 m
        if (b_mn) { for (int g = 0; g < BN / 64; ++g) tc::tma_load_2d(sB + g * 8192, &mapB, g * 64, 0, &bar_full); }
        else tc::tma_load_2d(sB, &mapB, 0, 0, &bar_full);
        tc::mbar_wait(&bar_full, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::umma_idesc_bf16_major(128, BN, a_mn != 0, b_mn != 0);
        for (int k = 0; k < 4; ++k) {       // K = 64 = 4 k-steps of 16
            const uint64_t ad = a_mn ? tc::umma_smem_desc_mn128(tc::smem_u32(sA) + k * 2048, 8192, 1024)
                                     : tc::umma_smem_desc_k128(tc::smem_u32(sA) + k * 32);
            const uint64_t bd = b_mn ? tc::umma_smem_desc_mn128(tc::smem_u32(sB) + k * 2048, 8192, 1024)
                                     : tc::umma_smem_desc_k128(tc::smem_u32(sB) + k * 32);
            tc::umma_bf16_ss(tmem_base, ad, bd, idesc, k ? 1u : 0u);
        }
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        float* o = out + (size_t)(warp * 32 + lane) * BN + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c + i < BN) o[i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

// CTA-pair form (cta_group::2): out[256 x BN] = A[256 x 128] * B[BN x 128]^T on a cluster of two CTAs.  CTA r loads A rows
// [128 r, 128 r + 128) and B rows [BN/2 r, BN/2 r + BN/2) into its own smem (the TMA credits the leader's mbarrier), the
// leader issues M = 256 MMAs, the commit is multicast to both CTAs, and each CTA reads its 128 accumulator rows from its
// own TMEM.  Pins the pair plumbing lp_topk_pair_kernel relies on.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_pair_selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int BN, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int half = BN / 2;
    uint8_t* sA = smem;                     // [2 kblocks][128 rows][128 B]
    uint8_t* sB = smem + 2 * 128 * 128;     // [2 kblocks][BN/2 rows][128 B]
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();

    if (warp == 0) tc::tmem_alloc_pair<256>(&tmem_base_s);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    tc::cluster_sync();                     // both CTAs' barriers exist before any remote traffic
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t leader_full = tc::mapa_u32(tc::smem_u32(&bar_full), 0);
        if (rank == 0) tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)(2 * (128 + half) * 256));    // bytes of BOTH CTAs
        for (int kb = 0; kb < 2; ++kb) {
            tc::tma_load_2d_pair(sA + kb * 128 * 128, &mapA, kb * 64, (int)rank * 128, leader_full);
            tc::tma_load_2d_pair(sB + kb * half * 128, &mapB, kb * 64, (int)rank * half, leader_full);
        }
        if (rank == 0) {
            tc::mbar_wait(&bar_full, 0);
            tc::tc_fence_after();
            const uint32_t idesc = tc::umma_idesc_bf16(256, BN);
            for (int kb = 0; kb < 2; ++kb)
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = tc::umma_smem_desc_k128(tc::smem_u32(sA + kb * 128 * 128) + k * 32);
                    const uint64_t bd = tc::umma_smem_desc_k128(tc::smem_u32(sB + kb * half * 128) + k * 32);
                    tc::umma_bf16_ss_pair(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
            tc::umma_commit_pair(&bar_mma, 0b11);
        }
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        float* o = out + (size_t)(rank * 128 + warp * 32 + lane) * BN + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c + i < BN) o[i] = v[i];
    }
    tc::tc_fence_before();
    tc::cluster_sync();                     // neither CTA may exit (or free TMEM) while the peer still uses the pair
    if (warp == 0) tc::tmem_dealloc_pair<256>(tmem_base);
}

}  // namespace crw

using namespace crw;

extern "C" int crw_debug_umma_pair_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream) {
    if (!A_bf16 || !B_bf16 || !out || BN < 32 || BN > 256 || (BN % 32)) return CRW_ERR_INVALID;
    CUtensorMap mA, mB;
    int rc = make_tmap_bf16_k64(&mA, A_bf16, 256, 128, 128);
    if (rc != CRW_OK) return rc;
    rc = make_tmap_bf16_k64(&mB, B_bf16, (uint64_t)BN, 128, (uint32_t)(BN / 2));
    if (rc != CRW_OK) return rc;
    const size_t smem = 1024 + 2 * 128 * 128 + (size_t)BN * 128;
    CRW_CUDA_RET(cudaFuncSetAttribute(umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pair_selftest_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(mA, mB, BN, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_debug_umma_mn_gemm(const void* A_bf16, const void* B_bf16, int BN, int a_mn, int b_mn, float* out, void* stream) {
    if (!A_bf16 || !B_bf16 || !out || (BN != 64 && BN != 128)) return CRW_ERR_INVALID;
    CUtensorMap mA, mB;
    // K-major: matrix [rows][64 k], box 64 k x rows.  MN-major: matrix [64 k][mn], box 64 mn x 64 k.
    int rc = a_mn ? make_tmap_bf16_box(&mA, A_bf16, 64, 128, 64, 64) : make_tmap_bf16_box(&mA, A_bf16, 128, 64, 128, 64);
    if (rc != CRW_OK) return rc;
    rc = b_mn ? make_tmap_bf16_box(&mB, B_bf16, 64, (uint64_t)BN, 64, 64) : make_tmap_bf16_box(&mB, B_bf16, (uint64_t)BN, 64, (uint32_t)BN, 64);
    if (rc != CRW_OK) return rc;
    const size_t smem = 1024 + 16384 + (size_t)BN * 128;
    CRW_CUDA_RET(cudaFuncSetAttribute(umma_mn_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_mn_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA, mB, BN, a_mn, b_mn, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_debug_umma_ts_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream) {
    if (!A_bf16 || !B_bf16 || !out || BN < 16 || BN > 256 || (BN % 16)) return CRW_ERR_INVALID;
    CUtensorMap mB;
    int rc = make_tmap_bf16_k64(&mB, B_bf16, (uint64_t)BN, 128, (uint32_t)BN);
    if (rc != CRW_OK) return rc;
    const size_t smem = 1024 + 2 * (size_t)BN * 128;
    CRW_CUDA_RET(cudaFuncSetAttribute(umma_ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_ts_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(A_bf16), mB, BN, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_debug_umma_tscp_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream) {
    if (!A_bf16 || !B_bf16 || !out || BN < 32 || BN > 256 || (BN % 32)) return CRW_ERR_INVALID;
    CUtensorMap mA, mB;
    int rc = make_tmap_bf16_k64(&mA, A_bf16, 128, 128, 128);
    if (rc != CRW_OK) return rc;
    rc = make_tmap_bf16_k64(&mB, B_bf16, (uint64_t)BN, 128, (uint32_t)BN);
    if (rc != CRW_OK) return rc;
    const size_t smem = 1024 + 2 * 128 * 128 + 2 * (size_t)BN * 128;
    CRW_CUDA_RET(cudaFuncSetAttribute(umma_tscp_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_tscp_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA, mB, BN, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_debug_umma_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream) {
    if (!A_bf16 || !B_bf16 || !out || BN < 16 || BN > 256 || (BN % 16)) return CRW_ERR_INVALID;
    CUtensorMap mA, mB;
    int rc = make_tmap_bf16_k64(&mA, A_bf16, 128, 128, 128);
    if (rc != CRW_OK) return rc;
    rc = make_tmap_bf16_k64(&mB, B_bf16, (uint64_t)BN, 128, (uint32_t)BN);
    if (rc != CRW_OK) return rc;
    const size_t smem = 1024 + 2 * 128 * 128 + 2 * (size_t)BN * 128;
    CRW_CUDA_RET(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA, mB, BN, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
