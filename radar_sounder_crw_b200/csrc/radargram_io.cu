// radargram_io.cu -- the data formats either side of the hot path, all HBM-bound copies:
//   * crw_patch_unfold     radargram [H, W] -> frame sequences [R, T, N, h, w]      src/dataset.py:34-39, src/utils.py:108
//   * crw_seed_labels      first-column reference labels + one-hot mask              src/utils.py:139-147
//   * crw_fuse_reversed    reversed-pass mask fusion of the two predicted maps       scripts/test/test_all.py:128-158
// Integer / index work: bit-exact against oracle/radargram_oracle.py.
#include "common.cuh"

namespace crw {

struct UnfoldGeom {
    int H, R, T, N, h, w, sh, sw, reverse;     // sh = h - oh, sw = w - ow (patch steps)
    long long ld, col_start, col_stride;
};

// out[r,t,n,y,x] = rg[n*sh + y][col_start + r*col_stride + t'*sw + x],  t' = reverse ? T-1-t : t
// One thread per VEC consecutive x; the output is written linearly (fully coalesced), the reads are w-element runs.
template <int VEC>
__global__ void __launch_bounds__(256) patch_unfold_kernel(const float* __restrict__ rg, UnfoldGeom g, float* __restrict__ out) {
    const int wv = g.w / VEC;
    const long long total = (long long)g.R * g.T * g.N * g.h * wv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long q = i;
        const int xv = (int)(q % wv); q /= wv;
        const int y = (int)(q % g.h); q /= g.h;
        const int n = (int)(q % g.N); q /= g.N;
        const int t = (int)(q % g.T);
        const int r = (int)(q / g.T);
        const int ts = g.reverse ? g.T - 1 - t : t;
        const float* src = rg + (long long)(n * g.sh + y) * g.ld + g.col_start + (long long)r * g.col_stride + (long long)ts * g.sw + xv * VEC;
        if (VEC == 4) reinterpret_cast<float4*>(out)[i] = *reinterpret_cast<const float4*>(src);
        else out[i] = *src;
    }
}

// label0[r,i] = seg[min(floor(i * float(rows)/float(N)), rows-1)][col_start + r*col_stride];  mask0[r,m,i] = (label0 == m)
__global__ void __launch_bounds__(256) seed_labels_kernel(const float* __restrict__ seg, int rows, long long ld, long long col_start,
                                                          long long col_stride, int R, int N, int M, int32_t* __restrict__ label0,
                                                          float* __restrict__ mask0) {
    const float scale = (float)rows / (float)N;
    const int total = R * N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int r = e / N, i = e % N;
        const int src = min((int)floorf((float)i * scale), rows - 1);
        const int lab = (int)seg[(long long)src * ld + col_start + (long long)r * col_stride];
        if (label0) label0[e] = lab;
        if (mask0)
            for (int m = 0; m < M; ++m) mask0[((size_t)r * M + m) * N + i] = (lab == m) ? 1.0f : 0.0f;
    }
}

// colok[x] = all_y rev[y][x] != 4   (test_all.py:152, in the reversed pass's own column order)
__global__ void __launch_bounds__(256) fuse_colok_kernel(const float* __restrict__ rev, int H, long long W, uint8_t* __restrict__ colok) {
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < W; x += (long long)gridDim.x * blockDim.x) {
        bool ok = true;
        for (int y = 0; y < H; ++y) ok = ok && (rev[(long long)y * W + x] != 4.0f);
        colok[x] = ok ? 1 : 0;
    }
}
// out[y][x] = mask ? 2 : fwd[y][x],  mask from the reversed pass un-flipped per radargram (test_all.py:146-158)
__global__ void __launch_bounds__(256) fuse_reversed_kernel(const float* __restrict__ fwd, const float* __restrict__ rev, int H, long long W,
                                                            int rg_len, int rule, const uint8_t* __restrict__ colok, float* __restrict__ out) {
    const long long total = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long x = i % W, y = i / W;
        const long long xr = (x / rg_len) * rg_len + (rg_len - 1 - x % rg_len);      // column of the reversed pass
        const float f = fwd[i];
        bool m = rev[y * W + xr] == 2.0f;
        if (rule == 1) m = m && (f != 3.0f) && colok[xr];
        if (rule == 3) m = m && (i >= total / 2);
        out[i] = m ? 2.0f : f;
    }
}

static inline unsigned grid_for(long long work_items) {
    const long long blocks = (work_items + 255) / 256;
    return (unsigned)(blocks < 148 * 16 ? (blocks < 1 ? 1 : blocks) : 148 * 16);
}

}  // namespace crw

using namespace crw;

extern "C" int crw_patch_unfold(const float* rg, int H, int64_t ld, int64_t col_start, int64_t col_stride, int R, int T, int N, int h,
                                int w, int oh, int ow, int reverse, float* out, void* stream) {
    if (H < 1 || ld < 1 || R < 0 || T < 0 || N < 1 || h < 1 || w < 1 || oh < 0 || ow < 0 || oh >= h || ow >= w || col_start < 0 ||
        col_stride < 0)
        return CRW_ERR_INVALID;
    if (R == 0 || T == 0) return CRW_OK;
    if (!rg || !out) return CRW_ERR_INVALID;
    const long long pxh = (long long)N * h - (long long)oh * (N - 1), pxw = (long long)T * w - (long long)ow * (T - 1);
    if (pxh > H || col_start + (long long)(R - 1) * col_stride + pxw > ld) return CRW_ERR_INVALID;     // the slice of dataset.py:35
    UnfoldGeom g{H, R, T, N, h, w, h - oh, w - ow, reverse ? 1 : 0, ld, col_start, col_stride};
    const bool vec = (w % 4 == 0) && (g.sw % 4 == 0) && (ld % 4 == 0) && (col_start % 4 == 0) && (col_stride % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(rg) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const long long total = (long long)R * T * N * h * (vec ? w / 4 : w);
    if (vec) patch_unfold_kernel<4><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(rg, g, out);
    else patch_unfold_kernel<1><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(rg, g, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_seed_labels(const float* seg, int rows, int64_t ld, int64_t col_start, int64_t col_stride, int R, int N, int M,
                               int32_t* label0, float* mask0, void* stream) {
    if (!seg || (!label0 && !mask0) || rows < 1 || ld < 1 || R < 0 || N < 1 || M < 1 || col_start < 0 || col_stride < 0) return CRW_ERR_INVALID;
    if (R == 0) return CRW_OK;
    if (col_start + (long long)(R - 1) * col_stride >= ld) return CRW_ERR_INVALID;
    seed_labels_kernel<<<grid_for((long long)R * N), 256, 0, (cudaStream_t)stream>>>(seg, rows, ld, col_start, col_stride, R, N, M, label0, mask0);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" size_t crw_fuse_reversed_scratch_bytes(int64_t W) { return W > 0 ? (size_t)W : 0; }

extern "C" int crw_fuse_reversed(const float* fwd, const float* rev, int H, int64_t W, int rg_len, int rule, float* out, void* scratch,
                                 size_t scratch_bytes, void* stream) {
    if (!fwd || !rev || !out || H < 1 || W < 0 || rg_len < 1) return CRW_ERR_INVALID;
    if (rule != 0 && rule != 1 && rule != 3) return CRW_ERR_UNSUPPORTED;
    if (W % rg_len) return CRW_ERR_INVALID;                 // test_all.py:77 trims seg to whole radargrams
    if (W == 0) return CRW_OK;
    uint8_t* colok = nullptr;
    if (rule == 1) {
        if (!scratch || scratch_bytes < (size_t)W) return CRW_ERR_WORKSPACE;
        colok = static_cast<uint8_t*>(scratch);
        fuse_colok_kernel<<<grid_for(W), 256, 0, (cudaStream_t)stream>>>(rev, H, W, colok);
        CRW_LAUNCH_RET();
    }
    fuse_reversed_kernel<<<grid_for((long long)H * W), 256, 0, (cudaStream_t)stream>>>(fwd, rev, H, W, rg_len, rule, colok, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}
