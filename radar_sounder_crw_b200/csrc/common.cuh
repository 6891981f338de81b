// common.cuh -- shared device helpers for libcrw_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/crw_b200.h"

#define CRW_CUDA_RET(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return CRW_ERR_CUDA_BASE - (int)_e; \
    } while (0)
#define CRW_LAUNCH_RET()                                     \
    do {                                                     \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) return CRW_ERR_CUDA_BASE - (int)_e; \
    } while (0)

namespace crw {

constexpr float kMaskBias = -1e10f;  // labelprop.py:94
constexpr float kNormEps = 1e-12f;   // F.normalize eps (model.py:22)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Pinned exp for x <= 0 -- bit-identical to crw_oracle_expf (oracle/crw_oracle.c): every
// step is one IEEE fp32 operation, written with intrinsics so nvcc cannot re-associate or fuse.
__device__ __forceinline__ float pinned_expf(float x) {
    if (x < -87.0f) return 0.0f;
    float t = __fmul_rn(x, 1.44269504088896341f);
    float n = rintf(t);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float z = __fmul_rn(r, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float y = __fmaf_rn(p, z, r);
    y = __fadd_rn(y, 1.0f);
    return __fmul_rn(y, __int_as_float(((int)n + 127) << 23));
}

__device__ __forceinline__ float warp_sum_butterfly_rn(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// trimmed key-frame rule of maskedatt.py:166-167
__host__ __device__ inline int n_key_frames(int n, int ctx) { return n <= ctx + 1 ? n : ctx + 1; }
__host__ __device__ inline int key_frame(int n, int ctx, int f) {
    if (n <= ctx + 1) return f;
    return f == 0 ? 0 : n - ctx + (f - 1);
}
// frame whose soft mask slot f gathers from (labelprop.py:82,106)
__host__ __device__ inline int label_frame(int n, int ctx, int f, int mode_fixed) {
    if (n <= ctx + 1 || mode_fixed) return key_frame(n, ctx, f);
    return f;
}

}  // namespace crw
