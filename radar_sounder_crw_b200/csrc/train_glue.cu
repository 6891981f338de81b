// train_glue.cu -- the train-loop glue of scripts/train.py:56,70-72: torch.optim.Adam over ALL encoder
// parameters as ONE elementwise launch on flat buffers (parameters, gradients, both moments: 16 bytes read and 12 written
// per parameter -- HBM-bound), and the zeroing of the flat gradient buffer that replaces optimizer.zero_grad().
// Algorithm = torch.optim.Adam (amsgrad = False, maximize = False), restated in oracle/adam_oracle.py:
//   g' = g + wd p;  m = m + (g' - m)(1 - b1);  v = b2 v + (1 - b2) g' g';
//   p = p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <algorithm>
#include <cmath>
#include "common.cuh"

namespace crw {

struct AdamHyper { float lr_over_bc1, inv_sqrt_bc2, omb1, beta2, omb2, eps, wd, gscale; };      // omb = 1 - beta, rounded from double

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHyper& h) {
    g = fmaf(h.wd, p, g * h.gscale);
    m = fmaf(g - m, h.omb1, m);
    v = fmaf(h.omb2, g * g, h.beta2 * v);
    const float denom = fmaf(sqrtf(v), h.inv_sqrt_bc2, h.eps);
    p = p - h.lr_over_bc1 * (m / denom);
}

// grid-stride over float4s; the scalar tail (n % 4) is handled by the first threads
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, AdamHyper h) {
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = tid; i < n4; i += stride) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, h);
        adam_one(pp.y, gg.y, mm.y, vv.y, h);
        adam_one(pp.z, gg.z, mm.z, vv.z, h);
        adam_one(pp.w, gg.w, mm.w, vv.w, h);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    const long long t = (n4 << 2) + tid;
    if (t < n) adam_one(p[t], g[t], m[t], v[t], h);
}

}  // namespace crw

using namespace crw;

extern "C" int crw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int64_t step, double grad_scale, void* stream) {
    if (n < 0 || step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return CRW_ERR_INVALID;
    if (n == 0) return CRW_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return CRW_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u)
        return CRW_ERR_ALIGN;
    // hyper-parameters are doubles, as torch.optim.Adam holds them (python floats): 1 - beta and the bias corrections are
    // formed in double and rounded once (1.0f - 0.999f is 1.3e-5 away from float(0.001))
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamHyper h;
    h.lr_over_bc1 = (float)(lr / bc1);
    h.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    h.omb1 = (float)(1.0 - beta1); h.beta2 = (float)beta2; h.omb2 = (float)(1.0 - beta2);
    h.eps = (float)eps; h.wd = (float)weight_decay; h.gscale = (float)grad_scale;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = ((n >> 2) + 255) / 256 + 1;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 16);
    adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, (long long)n, h);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
