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

// ---- the pinned dot product (oracle/crw_oracle.c::crw_oracle_dot) in its native form: one float4 of channels per lane ----
// The oracle's dot product (crw_oracle_dot) of a key row with the query row, both held one float4 per lane: fmaf chain over the
// lane's four channels, then the xor butterfly 16, 8, 4, 2, 1.  Every lane returns the same value.
__device__ __forceinline__ float x_warp_dot(const float4& kv, const float4& qv) {
    float acc = __fmaf_rn(kv.x, qv.x, 0.0f);
    acc = __fmaf_rn(kv.y, qv.y, acc);
    acc = __fmaf_rn(kv.z, qv.z, acc);
    acc = __fmaf_rn(kv.w, qv.w, acc);
    return warp_sum_butterfly_rn(acc);
}

// Sixteen warp dots at once.  p[v] = this lane's partial (fmaf chain over its four channels) of dot product v.  The xor butterfly
// of x_warp_dot is run as a REDUCE-SCATTER: at the level with offset 16 a lane keeps the eight vectors whose index bit 3 equals
// its lane bit 4 and hands the other eight to its partner, and so on (offsets 8, 4, 2), then the last level (offset 1) is a plain
// exchange.  Every sum adds the same two operands as the butterfly does at that level (fp addition is commutative), so the
// result is bit-identical to x_warp_dot -- with 16 shuffles instead of 80.  Returns the total of vector (lane >> 1).
__device__ __forceinline__ float x_warp_dot16(const float (&p)[16], int lane) {
    float q[8], r[4], s2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = b4 ? p[8 + i] : p[i], send = b4 ? p[i] : p[8 + i];
        q[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b3 ? q[4 + i] : q[i], send = b3 ? q[i] : q[4 + i];
        r[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b2 ? r[2 + i] : r[i], send = b2 ? r[i] : r[2 + i];
        s2[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4));
    }
    const float keep = b1 ? s2[1] : s2[0], send = b1 ? s2[0] : s2[1];
    const float t = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 2));
    return __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
}
__device__ __forceinline__ float x_partial4(const float4& kv, const float4& qv) {
    float acc = __fmaf_rn(kv.x, qv.x, 0.0f);
    acc = __fmaf_rn(kv.y, qv.y, acc);
    acc = __fmaf_rn(kv.z, qv.z, acc);
    return __fmaf_rn(kv.w, qv.w, acc);
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
