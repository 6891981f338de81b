// Profiling aid (standalone): cycles of the pieces of the warp-level bf16x3 tile used by walk_small.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kLD = 68;
template <int MODE>   // 0 full, 1 loads + split only, 2 mma only, 3 loads only
__global__ void probe(const float* in, float* out, long long* cyc, int K4) {
    __shared__ __align__(16) float A[64 * kLD], B[64 * kLD];
    for (int i = threadIdx.x; i < 64 * kLD; i += blockDim.x) { A[i] = in[i]; B[i] = in[i + 64 * kLD]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int m0 = 16 * (warp & 3), c0 = 32 * (warp >> 2);
    float acc[4][4] = {};
    uint32_t keep = 0;
    const long long t0 = clock64();
    for (int k0 = 0; k0 < K4; k0 += 16) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = m0 + g + 8 * r, k = k0 + 2 * tig + 8 * h;
                float x0 = 1.f, x1 = 2.f;
                if (MODE != 2) { const float2 v = *reinterpret_cast<const float2*>(A + m * kLD + k); x0 = v.x; x1 = v.y; }
                if (MODE == 3) { ah[2 * h + r] = __float_as_uint(x0); al[2 * h + r] = __float_as_uint(x1); }
                else split_pair(x0, x1, ah[2 * h + r], al[2 * h + r]);
            }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int n = c0 + 8 * nt + g;
            uint32_t bh[2], bl[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + 2 * tig + 8 * h;
                float x0 = 1.f, x1 = 2.f;
                if (MODE != 2) { x0 = B[k * kLD + n]; x1 = B[(k + 1) * kLD + n]; }
                if (MODE == 3) { bh[h] = __float_as_uint(x0); bl[h] = __float_as_uint(x1); }
                else split_pair(x0, x1, bh[h], bl[h]);
            }
            if (MODE == 0 || MODE == 2) {
                mma_bf16(acc[nt], ah, bh[0], bh[1]);
                mma_bf16(acc[nt], ah, bl[0], bl[1]);
                mma_bf16(acc[nt], al, bh[0], bh[1]);
            } else {
                keep ^= ah[0] ^ ah[1] ^ ah[2] ^ ah[3] ^ al[0] ^ al[1] ^ al[2] ^ al[3] ^ bh[0] ^ bh[1] ^ bl[0] ^ bl[1];
            }
        }
    }
    const long long t1 = clock64();
    float s = __uint_as_float(keep);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 2 * 64 * kLD * 4); cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 2 * 64 * kLD * 4);
    const char* names[4] = {"full", "loads+split", "mma only", "loads only"};
    for (int mode = 0; mode < 4; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) probe<0><<<1, 256>>>(in, out, cyc, 48);
            if (mode == 1) probe<1><<<1, 256>>>(in, out, cyc, 48);
            if (mode == 2) probe<2><<<1, 256>>>(in, out, cyc, 48);
            if (mode == 3) probe<3><<<1, 256>>>(in, out, cyc, 48);
            long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            if (rep == 1) printf("%-12s K=48, 256 threads: %lld cycles (%s)\n", names[mode], h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
