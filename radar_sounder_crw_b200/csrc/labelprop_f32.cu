// labelprop_f32.cu -- fp32 (pinned-order) label-propagation kernels for sm_100a.
//
// Replaces, for the h = N, w = 1 node grid the reference always uses:
//   F.normalize                     src/utils.py:115
//   batched_affinity                src/imported/maskedatt.py:151-175   (einsum, +mask, /temp, trim, topk, softmax)
//   radius mask                     src/imported/maskedatt.py:232-245, src/imported/labelprop.py:89-96
//   label gather of predict         src/imported/labelprop.py:82,106-109
//   frame loop + argmax             src/utils.py:152-160
//
// Arithmetic order is pinned (see oracle/crw_oracle.c header) so results are bit-identical
// to the C oracle: dot in WARP ORDER (lane l: fmaf chain over channels 4l .. 4l+3, then the xor butterfly 16, 8, 4, 2, 1 of
// plain adds -- crw_oracle_dot), logit = dot * (1/temp), ties by ascending candidate id, pinned polynomial exp, sequential
// softmax sum, mul-then-add gather.
#include "common.cuh"

namespace crw {

// ------------------------------------------------------------------------------------------
// L2 normalise: one warp per row.  lane l sums c = l, l+32, ... with fmaf, xor-butterfly.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* x, int64_t rows, int C, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* xr = x + row * C;
    float ss = 0.0f;
    for (int c = lane; c < C; c += 32) { float v = xr[c]; ss = __fmaf_rn(v, v, ss); }
    ss = warp_sum_butterfly_rn(ss);
    const float d = fmaxf(__fsqrt_rn(ss), kNormEps);
    float* o = out + row * C;
    for (int c = lane; c < C; c += 32) o[c] = __fdiv_rn(xr[c], d);
}

// ------------------------------------------------------------------------------------------
// Affinity + radius band + top-k + softmax, fp32.
//
// One CTA = one (query frame, chunk of 64 query nodes).  Query rows and key rows are staged row-major in shared memory
// ([64][128] floats, channels beyond C zero).  A (query, key) pair inside the band is one warp's work: every lane multiplies
// its float4 of channels (fmaf chain over 4l .. 4l+3) and the 32 partials meet in the xor butterfly -- the pinned order of
// crw_oracle_dot -- sixteen pairs at a time as a reduce-scatter (x_warp_dot16), which leaves each logit on a lane of its own;
// each warp keeps the running top-k of its queries sorted across lanes (ballot + shuffle insertion).  This is the validation /
// general-shape path (the exact tensor path of labelprop_x.cu is the fast one and gives the same bits).
// ------------------------------------------------------------------------------------------
constexpr int kChunk = 64;          // nodes per staged tile
constexpr int kScoreLd = 68;        // padded row of the score tile (16-byte aligned rows)
constexpr int kTopkThreads = 256;
constexpr int kQPerWarp = kChunk / (kTopkThreads / 32);  // 8 queries per warp
constexpr int kRowF = 128;          // floats per staged row (C <= 128)

struct TopkParams {
    const float* keys;     // [n_keys, N, C]
    const float* queries;  // [n_q, N, C]
    float* W;              // [n_q, k, N]
    int32_t* I;            // [n_q, k, N]
    int n_first, n_q, N, C, ctx, rb, k;
    float inv_temp;
};

// stage nodes [j_base, j_base+64) of `frame` ([N][C] row-major) into dst ([64][128], zero-padded)
__device__ __forceinline__ void stage_frame_rows(const float* __restrict__ frame, int N, int C, int j_base, float* dst) {
    const int c4n = C >> 2;
    for (int i = threadIdx.x; i < kChunk * (kRowF / 4); i += kTopkThreads) {
        const int jl = i >> 5, c4 = i & 31, j = j_base + jl;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < N && c4 < c4n) v = __ldg(reinterpret_cast<const float4*>(frame + (size_t)j * C) + c4);
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

template <int G>
__global__ void __launch_bounds__(kTopkThreads, 1) lp_topk_f32_kernel(TopkParams p) {
    extern __shared__ __align__(16) float smem[];
    const int N = p.N, C = p.C, k = p.k, rb = p.rb;
    float* Qr = smem;                              // [64][128]
    float* Kr = Qr + kChunk * kRowF;               // [G][64][128]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_qchunks = ceil_div(N, kChunk);
    const int job = blockIdx.x / n_qchunks, qc = blockIdx.x % n_qchunks;
    const int n = p.n_first + job;                 // the frame index this query plays
    const int F = n_key_frames(n, p.ctx);
    const int q_base = qc * kChunk;
    // key chunks that can intersect the band of this query chunk
    const int jc_lo = max(0, q_base - rb) / kChunk;
    const int jc_hi = min(N - 1, q_base + kChunk - 1 + rb) / kChunk;
    const int njc = jc_hi - jc_lo + 1;
    const int n_tiles = F * njc;

    stage_frame_rows(p.queries + (size_t)job * N * C, N, C, q_base, Qr);

    // running top-k of this warp's queries: lane i holds the i-th best (value, id)
    float tv[kQPerWarp];
    int ti[kQPerWarp];
#pragma unroll
    for (int i = 0; i < kQPerWarp; ++i) { tv[i] = -INFINITY; ti[i] = 0; }

    for (int t0 = 0; t0 < n_tiles; t0 += G) {
        const int g_cnt = min(G, n_tiles - t0);
        __syncthreads();   // previous products done with Kr
        for (int g = 0; g < g_cnt; ++g) {
            const int t = t0 + g, f = t / njc, jc = jc_lo + t % njc;
            stage_frame_rows(p.keys + (size_t)key_frame(n, p.ctx, f) * N * C, N, C, jc * kChunk, Kr + g * kChunk * kRowF);
        }
        __syncthreads();
        // dot products of the in-band pairs of this warp's queries in warp order (crw_oracle_dot), sixteen keys at a time
        // (x_warp_dot16: lane l ends up with the logit of key l >> 1), merged at once into the running top-k: candidates are
        // visited in ascending id order (tile order = frame-major, then node), so a later id never displaces an equal logit
        const unsigned kmask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
#pragma unroll 1
        for (int qi = 0; qi < kQPerWarp; ++qi) {
            const int ql = warp + qi * (kTopkThreads / 32);
            const int q = q_base + ql;
            if (q >= N) continue;                                     // warp-uniform
            const float4 qv = reinterpret_cast<const float4*>(Qr + ql * kRowF)[lane];
            float v = tv[qi];
            int id = ti[qi];
            float thr = __shfl_sync(0xffffffffu, v, k - 1);
            for (int g = 0; g < g_cnt; ++g) {
                const int t = t0 + g, f = t / njc, jc = jc_lo + t % njc;
                const float* Kg = Kr + g * kChunk * kRowF;
                const int j_lo = max(jc * kChunk, q - rb), j_hi = min(min(N - 1, jc * kChunk + kChunk - 1), q + rb);
                for (int j0 = j_lo; j0 <= j_hi; j0 += 16) {
                    float part[16];
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int jl = min(j0 + r, j_hi) - jc * kChunk;          // (rows past the band are recomputed and ignored)
                        part[r] = x_partial4(reinterpret_cast<const float4*>(Kg + jl * kRowF)[lane], qv);
                    }
                    const float cand = __fmul_rn(x_warp_dot16(part, lane), p.inv_temp);
                    const int j = j0 + (lane >> 1);
                    const bool ok = !(lane & 1) && j <= j_hi;
                    const int cid = f * N + j;
                    unsigned m = __ballot_sync(0xffffffffu, ok && cand > thr);
                    while (m) {
                        const int s = __ffs(m) - 1;
                        m &= m - 1;
                        const float c = __shfl_sync(0xffffffffu, cand, s);
                        const int ci = __shfl_sync(0xffffffffu, cid, s);
                        if (c > thr) {                  // warp-uniform
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
            }
            tv[qi] = v;
            ti[qi] = id;
        }
    }

    // finalise: masked fill (fewer than k in-band candidates), softmax, store
    const float masked = __fmul_rn(kMaskBias, p.inv_temp);
#pragma unroll
    for (int qi = 0; qi < kQPerWarp; ++qi) {
        const int q = q_base + warp + qi * (kTopkThreads / 32);
        if (q >= N) continue;
        float v = tv[qi];
        int id = ti[qi];
        int live = __popc(__ballot_sync(0xffffffffu, v > -INFINITY) & ((k >= 32) ? 0xffffffffu : ((1u << k) - 1u)));
        if (live < k) {
            // out-of-band candidates all carry the same logit; ascending id order
            for (int f = 0; f < F && live < k; ++f)
                for (int j = 0; j < N && live < k; ++j) {
                    const int dj = j - q;
                    if (dj <= rb && -dj <= rb) continue;
                    if (lane == live) { v = masked; id = f * N + j; }
                    ++live;
                }
        }
        const float v0 = __shfl_sync(0xffffffffu, v, 0);
        const float e = (lane < k) ? pinned_expf(__fsub_rn(v, v0)) : 0.0f;
        float s = __shfl_sync(0xffffffffu, e, 0);
        for (int j = 1; j < k; ++j) s = __fadd_rn(s, __shfl_sync(0xffffffffu, e, j));
        if (lane < k) {
            const size_t o = ((size_t)job * k + lane) * N + q;
            p.W[o] = __fdiv_rn(e, s);
            p.I[o] = id;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Label gather + argmax.
//   sequential kernel: one CTA per radargram walks frames [n_begin, n_end) in order (the
//   recurrence: frame n reads soft masks written for earlier frames).
//   parallel kernel: frames whose label frames are all final (ref_exact, n > ctx+1).
// `masks` is deliberately NOT __restrict__/ldg: it is read after being written by this CTA.
// ------------------------------------------------------------------------------------------
struct GatherParams {
    const float* W;      // [R,T,k,N]
    const int32_t* I;    // [R,T,k,N]
    const float* mask0;  // [R,M,N]
    int32_t* labels;     // [R,T,N]
    float* masks;        // [R,T,M,N]
    int R, T, N, M, ctx, k, mode_fixed;
};

__device__ __forceinline__ void gather_one(const GatherParams& p, const float* W, const int32_t* I, float* masks,
                                           int32_t* labels, int n, int q) {
    const int N = p.N, M = p.M, k = p.k;
    float best = 0.0f;
    int best_m = 0;
    for (int m0 = 0; m0 < M; m0 += 8) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
        for (int j = 0; j < k; ++j) {
            const size_t o = ((size_t)n * k + j) * N + q;
            const int id = I[o];
            const float w = W[o];
            const int lf = label_frame(n, p.ctx, id / N, p.mode_fixed);
            const float* src = masks + ((size_t)lf * M + m0) * N + (id % N);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (m0 + u < M) acc[u] = __fadd_rn(acc[u], __fmul_rn(src[(size_t)u * N], w));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (m0 + u < M) {
                masks[((size_t)n * M + m0 + u) * N + q] = acc[u];
                if ((m0 + u == 0) || acc[u] > best) { best = acc[u]; best_m = m0 + u; }
            }
    }
    labels[(size_t)n * N + q] = best_m;
}

__global__ void __launch_bounds__(256) lp_gather_seq_kernel(GatherParams p, int n_begin, int n_end, int init_frame0) {
    const int r = blockIdx.x;
    const int N = p.N, M = p.M;
    const float* W = p.W + (size_t)r * p.T * p.k * N;
    const int32_t* I = p.I + (size_t)r * p.T * p.k * N;
    float* masks = p.masks + (size_t)r * p.T * M * N;
    int32_t* labels = p.labels + (size_t)r * p.T * N;
    if (init_frame0) {
        const float* m0 = p.mask0 + (size_t)r * M * N;
        for (int q = threadIdx.x; q < N; q += blockDim.x) {
            float best = 0.0f;
            int bm = 0;
            for (int m = 0; m < M; ++m) {
                const float v = m0[(size_t)m * N + q];
                masks[(size_t)m * N + q] = v;
                if (m == 0 || v > best) { best = v; bm = m; }
            }
            labels[q] = bm;
        }
        __syncthreads();
    }
    for (int n = n_begin; n < n_end; ++n) {
        for (int q = threadIdx.x; q < N; q += blockDim.x) gather_one(p, W, I, masks, labels, n, q);
        __syncthreads();
    }
}

// Shared-memory version of the sequential gather: the soft masks of frame 0 and of the last ctx+1 frames
// live in a ring in shared memory (slot(f) = f == 0 ? 0 : 1 + (f-1) % (ctx+1); identical to f while
// f <= ctx+1), and each frame's W / I rows are prefetched with cp.async one frame ahead.
__device__ __forceinline__ int ring_slot(int f, int ctx) { return f == 0 ? 0 : 1 + (f - 1) % (ctx + 1); }

__global__ void __launch_bounds__(256) lp_gather_seq_smem_kernel(GatherParams p, int n_end) {
    extern __shared__ __align__(16) float gsm[];
    const int r = blockIdx.x, N = p.N, M = p.M, k = p.k, ctx = p.ctx;
    const int kn = k * N, mn = M * N;
    float* ring = gsm;                                     // [ctx+2][M][N]
    float* wbuf = ring + (size_t)(ctx + 2) * mn;           // [2][k*N]
    int* ibuf = reinterpret_cast<int*>(wbuf + 2 * kn);     // [2][k*N]
    const float* W = p.W + (size_t)r * p.T * kn;
    const int32_t* I = p.I + (size_t)r * p.T * kn;
    float* masks = p.masks + (size_t)r * p.T * mn;
    int32_t* labels = p.labels + (size_t)r * p.T * N;
    const float* m0 = p.mask0 + (size_t)r * mn;

    auto prefetch = [&](int n) {
        const uint32_t wdst = (uint32_t)__cvta_generic_to_shared(wbuf + (n & 1) * kn);
        const uint32_t idst = (uint32_t)__cvta_generic_to_shared(ibuf + (n & 1) * kn);
        for (int i = threadIdx.x; i < kn; i += blockDim.x) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wdst + 4 * i), "l"(W + (size_t)n * kn + i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(idst + 4 * i), "l"(I + (size_t)n * kn + i));
        }
        asm volatile("cp.async.commit_group;");
    };
    if (n_end > 1) prefetch(1);
    for (int i = threadIdx.x; i < mn; i += blockDim.x) { const float v = m0[i]; ring[i] = v; masks[i] = v; }
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        float best = 0.0f;
        int bm = 0;
        for (int m = 0; m < M; ++m) { const float v = m0[(size_t)m * N + q]; if (m == 0 || v > best) { best = v; bm = m; } }
        labels[q] = bm;
    }
    // three short phases per frame: (A) k*N threads form the products label * weight, (B) M*N threads add
    // them in top-k order (the pinned sequential sum), (C) N threads take the argmax.
    float* prod = reinterpret_cast<float*>(ibuf + 2 * kn);          // [k][M][N]
    const unsigned magic_n = (unsigned)(0x100000000ull / (unsigned)N) + 1u;
    int nm = 0;                                                      // (n-1) % (ctx+1)
    for (int n = 1; n < n_end; ++n) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                       // W/I of frame n landed; ring writes of frame n-1 visible
        if (n + 1 < n_end) prefetch(n + 1);
        const float* wn = wbuf + (n & 1) * kn;
        const int* in = ibuf + (n & 1) * kn;
        const bool trimmed = (n > ctx + 1);
        for (int idx = threadIdx.x; idx < kn; idx += blockDim.x) {
            const int j = (int)__umulhi((unsigned)idx, magic_n), q = idx - j * N;
            const int id = in[idx];
            const float w = wn[idx];
            const int f = (int)__umulhi((unsigned)id, magic_n), jj = id - f * N;
            // frame the slot f gathers from (labelprop.py:82,106) and its ring slot
            int slot;
            if (!trimmed) slot = f;                                   // frames 0..n-1, slot == frame
            else if (!p.mode_fixed) slot = f;                         // quirk: untrimmed list, frames 0..ctx
            else if (f == 0) slot = 0;
            else { const int d = ctx - f + 1; int s2 = nm - d; if (s2 < 0) s2 += ctx + 1; slot = 1 + s2; }
            const float* src = ring + (size_t)slot * mn + jj;
            float* dst = prod + (size_t)j * mn + q;
            for (int m = 0; m < M; ++m) dst[m * N] = __fmul_rn(src[m * N], w);
        }
        __syncthreads();
        float* out = ring + (size_t)(1 + nm) * mn;
        for (int idx = threadIdx.x; idx < mn; idx += blockDim.x) {
            float acc = 0.0f;
            for (int j = 0; j < k; ++j) acc = __fadd_rn(acc, prod[(size_t)j * mn + idx]);
            out[idx] = acc;
            masks[(size_t)n * mn + idx] = acc;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < N; q += blockDim.x) {
            float best = out[q];
            int best_m = 0;
            for (int m = 1; m < M; ++m) { const float v = out[(size_t)m * N + q]; if (v > best) { best = v; best_m = m; } }
            labels[(size_t)n * N + q] = best_m;
        }
        if (++nm == ctx + 1) nm = 0;
    }
}

__global__ void __launch_bounds__(256) lp_gather_par_kernel(GatherParams p, int n_begin, int n_end) {
    const int r = blockIdx.y;
    const int N = p.N;
    const float* W = p.W + (size_t)r * p.T * p.k * N;
    const int32_t* I = p.I + (size_t)r * p.T * p.k * N;
    float* masks = p.masks + (size_t)r * p.T * p.M * N;
    int32_t* labels = p.labels + (size_t)r * p.T * N;
    const long long total = (long long)(n_end - n_begin) * N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = n_begin + (int)(idx / N), q = (int)(idx % N);
        gather_one(p, W, I, masks, labels, n, q);
    }
}

// ref_exact tail (n > ctx+1): every id gathers from the frozen soft masks of frames 0..ctx, so they are
// staged once per CTA in shared memory; ids / weights of a query are fetched as one batch of independent loads.
__global__ void __launch_bounds__(256) lp_gather_par_smem_kernel(GatherParams p, int n_begin, int n_end) {
    extern __shared__ __align__(16) float lab[];      // [ctx+1][M][N]
    const int r = blockIdx.y, N = p.N, M = p.M, k = p.k;
    const int kn = k * N, mn = M * N;
    const float* W = p.W + (size_t)r * p.T * kn;
    const int32_t* I = p.I + (size_t)r * p.T * kn;
    float* masks = p.masks + (size_t)r * p.T * mn;
    int32_t* labels = p.labels + (size_t)r * p.T * N;
    for (int i = threadIdx.x; i < (p.ctx + 1) * mn; i += blockDim.x) lab[i] = masks[i];
    __syncthreads();
    const unsigned magic_n = (unsigned)(0x100000000ull / (unsigned)N) + 1u;
    const long long total = (long long)(n_end - n_begin) * N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = n_begin + (int)(idx / N), q = (int)(idx % N);
        const float* wq = W + (size_t)n * kn + q;
        const int32_t* iq = I + (size_t)n * kn + q;
        float best = 0.0f;
        int best_m = 0;
        for (int mb = 0; mb < M; mb += 8) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
            for (int j0 = 0; j0 < k; j0 += 8) {
                int ids[8];
                float ws[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = j0 + u < k;
                    ids[u] = ok ? __ldg(iq + (size_t)(j0 + u) * N) : 0;
                    ws[u] = ok ? __ldg(wq + (size_t)(j0 + u) * N) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (j0 + u < k) {
                        const int f = (int)__umulhi((unsigned)ids[u], magic_n), jj = ids[u] - f * N;
                        const float* src = lab + ((size_t)f * M + mb) * N + jj;   // slot f gathers from frame f (n > ctx+1)
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            if (mb + m < M) acc[m] = __fadd_rn(acc[m], __fmul_rn(src[m * N], ws[u]));
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (mb + m < M) {
                    masks[((size_t)n * M + mb + m) * N + q] = acc[m];
                    if ((mb + m == 0) || acc[m] > best) { best = acc[m]; best_m = mb + m; }
                }
        }
        labels[(size_t)n * N + q] = best_m;
    }
}

// ------------------------------------------------------------------------------------------
// Horizontality metric (src/utils.py:118-123): A[t][c][n] = <emb[t,c,:-1], emb[t,n,1:]> / 0.1,
// xent[n,t] = logsumexp_c A[t][c][n] - A[t][n][n].  One CTA per frame t < T-1, warp per column n.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xent_kernel(const float* __restrict__ emb, int T, int N, int C, float* xent) {
    extern __shared__ float fr[];  // [N][C+1]
    const int t = blockIdx.x, ld = C + 1;
    const float* e = emb + (size_t)t * N * C;
    for (int i = threadIdx.x; i < N * C; i += blockDim.x) fr[(i / C) * ld + (i % C)] = e[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < N; n += blockDim.x >> 5) {
        float mx = -INFINITY, se = 0.0f, diag = 0.0f;
        for (int c0 = 0; c0 < N; c0 += 32) {
            const int c = c0 + lane;
            float a = -INFINITY;
            if (c < N) {
                float acc = 0.0f;
                for (int ch = 0; ch < C - 1; ++ch) acc = fmaf(fr[c * ld + ch], fr[n * ld + ch + 1], acc);
                a = acc / 0.1f;
                if (c == n) diag = a;
            }
            const float cm = warp_max(a);
            const float nm = fmaxf(mx, cm);
            se = se * __expf(mx - nm) + warp_sum(c < N ? __expf(a - nm) : 0.0f);
            mx = nm;
        }
        diag = warp_sum(diag);
        if (lane == 0) xent[(size_t)n * (T - 1) + t] = logf(se) + mx - diag;
    }
}

// one stepwise predict call (labelprop.py:106-116): lbl [F,M,N] are the soft masks the ids index into
__global__ void __launch_bounds__(256) lp_gather_step_kernel(const float* __restrict__ W, const int32_t* __restrict__ I,
                                                             const float* __restrict__ lbl, int F, int N, int M, int k,
                                                             float* out_mask, int32_t* out_label) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    float best = 0.0f;
    int best_m = 0;
    for (int m = 0; m < M; ++m) {
        float acc = 0.0f;
        for (int j = 0; j < k; ++j) {
            const int id = I[(size_t)j * N + q];
            const int f = id / N;
            const float v = (f < F) ? lbl[((size_t)f * M + m) * N + (id % N)] : 0.0f;
            acc = __fadd_rn(acc, __fmul_rn(v, W[(size_t)j * N + q]));
        }
        out_mask[(size_t)m * N + q] = acc;
        if (m == 0 || acc > best) { best = acc; best_m = m; }
    }
    if (out_label) out_label[q] = best_m;
}

// Nearest-neighbour upsample of a label map (reference: `up = Resize((seg_h, rg_len), NEAREST)` applied to
// final_prediction[N,T], scripts/test/test_all.py:79,96).  labels [R,T,N] i32 -> out [R,H,W] f32 with torch's
// rule src = min(floor(dst * float(in/out)), in-1) on both axes.
__global__ void __launch_bounds__(256) lp_upsample_kernel(const int32_t* __restrict__ labels, int R, int T, int N, int H, int W,
                                                          float* __restrict__ out) {
    const float sy = (float)N / (float)H, sx = (float)T / (float)W;
    const long long total = (long long)R * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), r = (int)(i / ((long long)W * H));
        const int n = min((int)floorf((float)y * sy), N - 1), t = min((int)floorf((float)x * sx), T - 1);
        out[i] = (float)labels[((size_t)r * T + t) * N + n];
    }
}

}  // namespace crw

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace crw;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int crw_l2_normalize(const float* x, int64_t rows, int C, float* out, void* stream) {
    if (!x || !out || rows < 0 || C < 1) return CRW_ERR_INVALID;
    if (rows == 0) return CRW_OK;
    const int64_t blocks = (rows + 7) / 8;
    if (blocks > 0x7fffffffLL) return CRW_ERR_UNSUPPORTED;
    l2_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, C, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

template <int G>
static int launch_topk_f32(const TopkParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(1 + G) * kRowF * kChunk * sizeof(float);
    CRW_CUDA_RET(cudaFuncSetAttribute(lp_topk_f32_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.n_q * ceil_div(p.N, kChunk);
    lp_topk_f32_kernel<G><<<grid, kTopkThreads, smem, st>>>(p);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_affinity_topk(const float* keys, const float* queries, int n_first, int n_q, int N, int C, int ctx,
                                 float radius, float temp, int k, int precision, float* W, int32_t* I, void* stream) {
    if (!keys || !queries || !W || !I) return CRW_ERR_INVALID;
    if (n_first < 1 || n_q < 0 || N < 1 || C < 1 || ctx < 1 || k < 1 || !(radius > 0.0f) || !(temp > 0.0f))
        return CRW_ERR_INVALID;
    if ((int64_t)n_key_frames(n_first, ctx) * N < k) return CRW_ERR_INVALID;  // torch.topk would raise
    if (n_q == 0) return CRW_OK;
    if (k > 32) return CRW_ERR_UNSUPPORTED;
    if ((C & 3) || !aligned16(keys) || !aligned16(queries)) return CRW_ERR_ALIGN;
    if (precision == CRW_PREC_FP32) {
        if (C > 128) return CRW_ERR_UNSUPPORTED;
        TopkParams p;
        p.keys = keys; p.queries = queries; p.W = W; p.I = I;
        p.n_first = n_first; p.n_q = n_q; p.N = N; p.C = C; p.ctx = ctx; p.k = k;
        const float rc = ceilf(radius);
        p.rb = (rc - 1.0f >= (float)N) ? N : (int)rc - 1;   // |d| < radius  <=>  |d| <= ceil(radius)-1
        p.inv_temp = 1.0f / temp;
        return launch_topk_f32<3>(p, (cudaStream_t)stream);
    }
    return CRW_ERR_UNSUPPORTED;
}

// frames that must run in order: all of them in fixed mode, the first ctx+1 in ref_exact mode
static int gather_seq_end(const GatherParams& p) { return p.mode_fixed ? p.T : min(p.T, p.ctx + 2); }

static int gather_launch_seq(const GatherParams& p, cudaStream_t st) {
    const int seq_end = gather_seq_end(p);
    const size_t gsmem = ((size_t)(p.ctx + 2) * p.M * p.N + 4 * (size_t)p.k * p.N + (size_t)p.k * p.M * p.N) * sizeof(float);
    if (gsmem <= 200 * 1024) {
        if (gsmem > 48 * 1024)
            CRW_CUDA_RET(cudaFuncSetAttribute(lp_gather_seq_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        lp_gather_seq_smem_kernel<<<p.R, 256, gsmem, st>>>(p, seq_end);
    } else {
        lp_gather_seq_kernel<<<p.R, 256, 0, st>>>(p, 1, seq_end, 1);
    }
    CRW_LAUNCH_RET();
    return CRW_OK;
}

static int gather_launch_par(const GatherParams& p, cudaStream_t st) {
    const int seq_end = gather_seq_end(p), T = p.T;
    if (seq_end >= T) return CRW_OK;
    const long long total = (long long)(T - seq_end) * p.N;
    long long gx64 = (total + 255) / 256; int gx = (int)(gx64 < 148LL * 8 ? gx64 : 148LL * 8);
    const size_t psmem = (size_t)(p.ctx + 1) * p.M * p.N * sizeof(float);
    if (psmem <= 96 * 1024) {     // ref_exact only reaches here (fixed mode is sequential throughout)
        if (psmem > 48 * 1024)
            CRW_CUDA_RET(cudaFuncSetAttribute(lp_gather_par_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        if (gx > 148 * 2) gx = 148 * 2;
        lp_gather_par_smem_kernel<<<dim3(gx, p.R), 256, psmem, st>>>(p, seq_end, T);
    } else {
        lp_gather_par_kernel<<<dim3(gx, p.R), 256, 0, st>>>(p, seq_end, T);
    }
    CRW_LAUNCH_RET();
    return CRW_OK;
}

static int gather_params(GatherParams& p, const float* W, const int32_t* I, const float* mask0, int R, int T, int N, int M, int ctx,
                         int k, int mode, int32_t* labels, float* masks) {
    if (!W || !I || !mask0 || !labels || !masks) return CRW_ERR_INVALID;
    if (R < 0 || T < 1 || N < 1 || M < 1 || ctx < 1 || k < 1) return CRW_ERR_INVALID;
    if (mode != CRW_LP_REF_EXACT && mode != CRW_LP_FIXED) return CRW_ERR_INVALID;
    p.W = W; p.I = I; p.mask0 = mask0; p.labels = labels; p.masks = masks;
    p.R = R; p.T = T; p.N = N; p.M = M; p.ctx = ctx; p.k = k; p.mode_fixed = (mode == CRW_LP_FIXED);
    return CRW_OK;
}

extern "C" int crw_label_gather(const float* W, const int32_t* I, const float* mask0, int R, int T, int N, int M,
                                int ctx, int k, int mode, int32_t* labels, float* masks, void* stream) {
    GatherParams p;
    int rc = gather_params(p, W, I, mask0, R, T, N, M, ctx, k, mode, labels, masks);
    if (rc != CRW_OK) return rc;
    if (R == 0) return CRW_OK;
    if ((rc = gather_launch_seq(p, (cudaStream_t)stream)) != CRW_OK) return rc;
    return gather_launch_par(p, (cudaStream_t)stream);
}

extern "C" int crw_label_gather_step(const float* W, const int32_t* I, const float* lbl, int F, int N, int M, int k,
                                     float* out_mask, int32_t* out_label_or_null, void* stream) {
    if (!W || !I || !lbl || !out_mask || F < 1 || N < 1 || M < 1 || k < 1) return CRW_ERR_INVALID;
    lp_gather_step_kernel<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(W, I, lbl, F, N, M, k, out_mask,
                                                                            out_label_or_null);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

namespace crw {
// tensor path (labelprop_tc.cu): prep kernel + plan, then top-k launches over ranges of schedule slots
struct LpTcPlanStorage { alignas(64) unsigned char bytes[1024]; };
size_t lp_tc_split_bytes();
int lp_tc_prepare(const float* feats, int R, int T, int N, int C, int ctx, float radius, float temp, int k, int do_normalize,
                  float* W, int32_t* I, void* scratch, void* split_ws, cudaStream_t st, void* plan_storage);
int lp_tc_launch(const void* plan_storage, int v_begin, int v_end, int max_ctas, cudaStream_t st, bool split_tail);
int lp_tc_total_slots(const void* plan_storage);
int lp_tc_early_slots(const void* plan_storage);
int lp_tc_prep_rows(const float* stage, int64_t row_begin, int64_t nrows, int64_t total_rows, int do_normalize, void* scratch,
                    cudaStream_t st);
int lp_tc_launch_tiles(const void* plan_storage, int rg, int ta, int tb, int max_ctas, cudaStream_t st);
int lp_tc_tiles_per_rg(const void* plan_storage);
int lp_tc_tile_rows(const void* plan_storage);
// exact tensor path (labelprop_x.cu): prep + plan, then filter + refine launches over ranges of schedule slots
struct LpXPlanStorage { alignas(64) unsigned char bytes[1024]; };
size_t lp_x_plan_bytes();
size_t lp_x_scratch_bytes(int R, int T, int N, int C, int k, int do_normalize);
int lp_x_max_k();
int lp_x_prepare(const float* feats, int R, int T, int N, int C, int ctx, float radius, float temp, int k, int do_normalize, float* W,
                 int32_t* I, void* scratch, int sms, cudaStream_t st, void* plan_storage, size_t plan_bytes, int n_min = 1);
int lp_x_launch(const void* plan_storage, int v_begin, int v_end, int max_ctas, cudaStream_t st);
int lp_x_total_slots(const void* plan_storage);
int lp_x_early_slots(const void* plan_storage);
}

// A second, higher-priority stream and two events per device, created on first use: the tensor path forks the early query
// tiles + the sequential label gather onto it so that they overlap the bulk of the top-k (fork / join by events only,
// nothing synchronises with the host; the enqueue sequence is serialised by a mutex because the events are shared).
#include <mutex>
namespace {
struct SideCtx {
    cudaStream_t s2 = nullptr, s_copy = nullptr;                 // side compute stream (high priority), H2D copy stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};   // staging double buffer
    bool ok = false;
    std::mutex mu;
};
// environment switches are read once per process
bool env_no_fork() {
    static const bool v = [] { const char* e = getenv("CRW_LP_NO_FORK"); return e && atoi(e) != 0; }();
    return v;
}
SideCtx* side_ctx() {
    static SideCtx ctx[64];
    static std::once_flag once[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::call_once(once[dev], [dev]() {
        SideCtx& c = ctx[dev];
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);          // hi = numerically lowest = greatest priority
        c.ok = cudaStreamCreateWithPriority(&c.s2, cudaStreamNonBlocking, hi) == cudaSuccess &&
               cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming) == cudaSuccess &&
               cudaStreamCreateWithFlags(&c.s_copy, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2 && c.ok; ++i)
            c.ok = cudaEventCreateWithFlags(&c.ev_copied[i], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&c.ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    });
    return ctx[dev].ok ? &ctx[dev] : nullptr;
}
}  // namespace

extern "C" size_t crw_labelprop_scratch_bytes(int R, int T, int N, int C, int k, int precision, int do_normalize,
                                              int have_topk_out) {
    size_t b = 0;
    if (precision == CRW_PREC_BF16X3) b += align_up((size_t)R * T * N * C * 2 * 2, 256) + align_up(crw::lp_tc_split_bytes(), 256);   // bf16 hi + lo, tail-split lists
    else if (precision == CRW_PREC_TC_EXACT) b += align_up(crw::lp_x_scratch_bytes(R, T, N, C, k, do_normalize), 256);
    else if (do_normalize) b += align_up((size_t)R * T * N * C * sizeof(float), 256);
    if (!have_topk_out) b += 2 * align_up((size_t)R * T * k * N * sizeof(float), 256);
    return b + 256;
}

extern "C" int crw_labelprop_forward(const float* feats, const float* mask0, int R, int T, int N, int C, int M, int ctx,
                                     float radius, float temp, int k, int mode, int precision, int do_normalize,
                                     int32_t* labels, float* masks, float* W_or_null, int32_t* I_or_null,
                                     void* scratch, size_t scratch_bytes, void* stream) {
    if (!feats || !mask0 || !labels || !masks) return CRW_ERR_INVALID;
    if (R < 0 || T < 1 || N < 1 || C < 1 || M < 1) return CRW_ERR_INVALID;
    if ((W_or_null == nullptr) != (I_or_null == nullptr)) return CRW_ERR_INVALID;
    if (R == 0) return CRW_OK;
    const size_t need = crw_labelprop_scratch_bytes(R, T, N, C, k, precision, do_normalize, W_or_null != nullptr);
    if (need > 256 && (!scratch || scratch_bytes < need)) return CRW_ERR_WORKSPACE;
    char* sp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
    if (precision == CRW_PREC_BF16X3) {
        if (ctx < 1 || k < 1 || !(radius > 0.0f) || !(temp > 0.0f) || (int64_t)N < k) return CRW_ERR_INVALID;
        void* hilo = sp;
        sp += align_up((size_t)R * T * N * C * 2 * 2, 256);
        void* split_ws = sp;
        sp += align_up(lp_tc_split_bytes(), 256);
        float* Wt = W_or_null;
        int32_t* It = I_or_null;
        if (!Wt) {
            Wt = reinterpret_cast<float*>(sp);
            sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
            It = reinterpret_cast<int32_t*>(sp);
        }
        cudaStream_t st = (cudaStream_t)stream;
        GatherParams gp;
        int rc = gather_params(gp, Wt, It, mask0, R, T, N, M, ctx, k, mode, labels, masks);
        if (rc != CRW_OK) return rc;
        LpTcPlanStorage plan;
        if ((rc = lp_tc_prepare(feats, R, T, N, C, ctx, radius, temp, k, do_normalize, Wt, It, hilo, split_ws, st, plan.bytes)) != CRW_OK) return rc;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int total = lp_tc_total_slots(plan.bytes), early = lp_tc_early_slots(plan.bytes);
        // fork only when there is something to overlap: ref_exact mode (later frames do not depend on each other), query
        // tiles beyond the early ones, and few enough early tiles that they are a side job
        SideCtx* sc = (gp.mode_fixed || early >= total || early > sms / 2 || env_no_fork()) ? nullptr : side_ctx();
        if (!sc) {
            if ((rc = lp_tc_launch(plan.bytes, 0, total, sms, st, true)) != CRW_OK) return rc;
            if ((rc = gather_launch_seq(gp, st)) != CRW_OK) return rc;
            return gather_launch_par(gp, st);
        }
        std::lock_guard<std::mutex> lock(sc->mu);
        CRW_CUDA_RET(cudaEventRecord(sc->ev_fork, st));                       // prep done
        // the bulk is the critical path: enqueue it first (it leaves the early tiles their SMs when they are few)
        const int bulk_ctas = (early <= sms / 8) ? sms - early : sms;
        if ((rc = lp_tc_launch(plan.bytes, early, total, bulk_ctas, st, true)) != CRW_OK) return rc;
        CRW_CUDA_RET(cudaStreamWaitEvent(sc->s2, sc->ev_fork, 0));
        if ((rc = lp_tc_launch(plan.bytes, 0, early, sms, sc->s2, false)) != CRW_OK) return rc;     // frames 1..ctx+1 (and a few more)
        if ((rc = gather_launch_seq(gp, sc->s2)) != CRW_OK) return rc;                              // the true recurrence
        CRW_CUDA_RET(cudaEventRecord(sc->ev_join, sc->s2));
        CRW_CUDA_RET(cudaStreamWaitEvent(st, sc->ev_join, 0));
        return gather_launch_par(gp, st);
    }
    if (precision == CRW_PREC_TC_EXACT) {
        if (ctx < 1 || k < 1 || !(radius > 0.0f) || !(temp > 0.0f) || (int64_t)N < k) return CRW_ERR_INVALID;
        void* xs = sp;
        sp += align_up(lp_x_scratch_bytes(R, T, N, C, k, do_normalize), 256);
        float* Wt = W_or_null;
        int32_t* It = I_or_null;
        if (!Wt) {
            Wt = reinterpret_cast<float*>(sp);
            sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
            It = reinterpret_cast<int32_t*>(sp);
        }
        cudaStream_t st = (cudaStream_t)stream;
        GatherParams gp;
        int rc = gather_params(gp, Wt, It, mask0, R, T, N, M, ctx, k, mode, labels, masks);
        if (rc != CRW_OK) return rc;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        LpXPlanStorage plan;
        if ((rc = lp_x_prepare(feats, R, T, N, C, ctx, radius, temp, k, do_normalize, Wt, It, xs, sms, st, plan.bytes, sizeof(plan.bytes))) != CRW_OK) return rc;
        const int total = lp_x_total_slots(plan.bytes), early = lp_x_early_slots(plan.bytes);
        // the frames 1..ctx+1 (the only true recurrence) go to the side stream: their items, then the sequential gather
        SideCtx* sc = (gp.mode_fixed || early >= total || early > sms / 2 || env_no_fork()) ? nullptr : side_ctx();
        if (!sc) {
            if ((rc = lp_x_launch(plan.bytes, 0, total, sms, st)) != CRW_OK) return rc;
            if ((rc = gather_launch_seq(gp, st)) != CRW_OK) return rc;
            return gather_launch_par(gp, st);
        }
        std::lock_guard<std::mutex> lock(sc->mu);
        CRW_CUDA_RET(cudaEventRecord(sc->ev_fork, st));                       // prep done
        // the bulk is the critical path: enqueue it first (it leaves the early items their SMs when they are few)
        const int bulk_ctas = (early <= sms / 8) ? sms - early : sms;
        if ((rc = lp_x_launch(plan.bytes, early, total, bulk_ctas, st)) != CRW_OK) return rc;
        CRW_CUDA_RET(cudaStreamWaitEvent(sc->s2, sc->ev_fork, 0));
        if ((rc = lp_x_launch(plan.bytes, 0, early, sms, sc->s2)) != CRW_OK) return rc;
        if ((rc = gather_launch_seq(gp, sc->s2)) != CRW_OK) return rc;
        CRW_CUDA_RET(cudaEventRecord(sc->ev_join, sc->s2));
        CRW_CUDA_RET(cudaStreamWaitEvent(st, sc->ev_join, 0));
        return gather_launch_par(gp, st);
    }
    if (precision != CRW_PREC_FP32) return CRW_ERR_UNSUPPORTED;
    const float* emb = feats;
    if (do_normalize) {
        float* e = reinterpret_cast<float*>(sp);
        sp += align_up((size_t)R * T * N * C * sizeof(float), 256);
        int rc = crw_l2_normalize(feats, (int64_t)R * T * N, C, e, stream);
        if (rc != CRW_OK) return rc;
        emb = e;
    }
    float* W = W_or_null;
    int32_t* I = I_or_null;
    if (!W) {
        W = reinterpret_cast<float*>(sp);
        sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
        I = reinterpret_cast<int32_t*>(sp);
    }
    if (T > 1) {
        for (int r = 0; r < R; ++r) {
            const float* er = emb + (size_t)r * T * N * C;
            const size_t o = ((size_t)r * T + 1) * k * N;
            int rc = crw_affinity_topk(er, er + (size_t)N * C, 1, T - 1, N, C, ctx, radius, temp, k, precision, W + o,
                                       I + o, stream);
            if (rc != CRW_OK) return rc;
        }
    }
    return crw_label_gather(W, I, mask0, R, T, N, M, ctx, k, mode, labels, masks, stream);
}

// ------------------------------------------------------------------------------------------
// Host-streamed tensor path: `feats` lives in PINNED HOST memory.  Each radargram is cut into chunks of whole query tiles;
// chunk c+1 is copied into a staging double buffer on a copy stream while chunk c is normalised, split and searched
// (a query tile only needs key rows that precede it, so the top-k of a chunk can run as soon as its rows are resident).
// ------------------------------------------------------------------------------------------
static int host_chunk_tiles(int tiles_per_rg, int tile_rows, int C) {
    const size_t rg_bytes = (size_t)tiles_per_rg * tile_rows * C * sizeof(float);
    int nch = rg_bytes >= (size_t)24 << 20 ? 4 : (rg_bytes >= (size_t)8 << 20 ? 2 : 1);
    return ceil_div(tiles_per_rg, nch);
}

extern "C" size_t crw_labelprop_host_scratch_bytes(int R, int T, int N, int C, int k, int have_topk_out) {
    const int tile_rows = 256;                                    // upper bound over both top-k kernels
    const int tpr = ceil_div(T * N, 128);
    const size_t stage = align_up((size_t)(host_chunk_tiles(tpr, 128, C) * 128 + tile_rows) * C * sizeof(float), 256);
    return crw_labelprop_scratch_bytes(R, T, N, C, k, CRW_PREC_BF16X3, 1, have_topk_out) + 2 * stage;
}

extern "C" int crw_labelprop_forward_host(const float* feats_host, const float* mask0, int R, int T, int N, int C, int M, int ctx,
                                          float radius, float temp, int k, int mode, int do_normalize, int32_t* labels,
                                          float* masks, float* W_or_null, int32_t* I_or_null, void* scratch,
                                          size_t scratch_bytes, void* stream) {
    if (!feats_host || !mask0 || !labels || !masks) return CRW_ERR_INVALID;
    if (R < 0 || T < 1 || N < 1 || C < 1 || M < 1) return CRW_ERR_INVALID;
    if ((W_or_null == nullptr) != (I_or_null == nullptr)) return CRW_ERR_INVALID;
    if (ctx < 1 || k < 1 || !(radius > 0.0f) || !(temp > 0.0f) || (int64_t)N < k) return CRW_ERR_INVALID;
    if (R == 0) return CRW_OK;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, feats_host) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return CRW_ERR_INVALID;                                   // pageable memory would serialise the copies silently
    }
    const size_t need = crw_labelprop_host_scratch_bytes(R, T, N, C, k, W_or_null != nullptr);
    if (!scratch || scratch_bytes < need) return CRW_ERR_WORKSPACE;
    SideCtx* sc = side_ctx();
    if (!sc) return CRW_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    char* sp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
    void* hilo = sp;
    sp += align_up((size_t)R * T * N * C * 2 * 2, 256);
    float* Wt = W_or_null;
    int32_t* It = I_or_null;
    if (!Wt) {
        Wt = reinterpret_cast<float*>(sp);
        sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
        It = reinterpret_cast<int32_t*>(sp);
        sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
    }
    GatherParams gp;
    int rc = gather_params(gp, Wt, It, mask0, R, T, N, M, ctx, k, mode, labels, masks);
    if (rc != CRW_OK) return rc;
    LpTcPlanStorage plan;
    if ((rc = lp_tc_prepare(nullptr, R, T, N, C, ctx, radius, temp, k, do_normalize, Wt, It, hilo, nullptr, st, plan.bytes)) != CRW_OK) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tpr = lp_tc_tiles_per_rg(plan.bytes), tile_rows = lp_tc_tile_rows(plan.bytes);
    const int chunk_tiles = T < 2 ? 1 : host_chunk_tiles(tpr, tile_rows, C);
    const int64_t rows_rg = (int64_t)T * N, total_rows = (int64_t)R * rows_rg;
    const size_t stage_bytes = align_up((size_t)(host_chunk_tiles(ceil_div(T * N, 128), 128, C) * 128 + 256) * C * sizeof(float), 256);
    float* stage[2] = {reinterpret_cast<float*>(sp), reinterpret_cast<float*>(sp + stage_bytes)};

    std::lock_guard<std::mutex> lock(sc->mu);
    CRW_CUDA_RET(cudaEventRecord(sc->ev_fork, st));               // earlier work on `st` may still use the scratch
    CRW_CUDA_RET(cudaStreamWaitEvent(sc->s_copy, sc->ev_fork, 0));
    int g = 0;
    bool forked = false;
    for (int rg = 0; rg < R; ++rg) {
        for (int ta = 0; ta * (int64_t)tile_rows < rows_rg; ta += chunk_tiles, ++g) {
            const int slot = g & 1;
            const int64_t r0 = (int64_t)ta * tile_rows;
            const int64_t r1 = (r0 + (int64_t)chunk_tiles * tile_rows < rows_rg) ? r0 + (int64_t)chunk_tiles * tile_rows : rows_rg;
            const int64_t grow = rg * rows_rg + r0;
            if (g >= 2) CRW_CUDA_RET(cudaStreamWaitEvent(sc->s_copy, sc->ev_free[slot], 0));
            CRW_CUDA_RET(cudaMemcpyAsync(stage[slot], feats_host + grow * C, (size_t)(r1 - r0) * C * sizeof(float),
                                         cudaMemcpyHostToDevice, sc->s_copy));
            CRW_CUDA_RET(cudaEventRecord(sc->ev_copied[slot], sc->s_copy));
            CRW_CUDA_RET(cudaStreamWaitEvent(st, sc->ev_copied[slot], 0));
            if ((rc = lp_tc_prep_rows(stage[slot], grow, r1 - r0, total_rows, do_normalize, hilo, st)) != CRW_OK) return rc;
            CRW_CUDA_RET(cudaEventRecord(sc->ev_free[slot], st));
            if (T >= 2 && (rc = lp_tc_launch_tiles(plan.bytes, rg, ta, ta + chunk_tiles, sms, st)) != CRW_OK) return rc;
            // one radargram, first chunk holds frames 0..ctx+1: start the sequential gather beside the later chunks
            if (R == 1 && ta == 0 && !gp.mode_fixed && r1 < rows_rg && r1 >= (int64_t)(ctx + 2) * N) {
                CRW_CUDA_RET(cudaEventRecord(sc->ev_join, st));
                CRW_CUDA_RET(cudaStreamWaitEvent(sc->s2, sc->ev_join, 0));
                if ((rc = gather_launch_seq(gp, sc->s2)) != CRW_OK) return rc;
                CRW_CUDA_RET(cudaEventRecord(sc->ev_join, sc->s2));
                forked = true;
            }
        }
    }
    if (forked) CRW_CUDA_RET(cudaStreamWaitEvent(st, sc->ev_join, 0));
    else if ((rc = gather_launch_seq(gp, st)) != CRW_OK) return rc;
    return gather_launch_par(gp, st);
}

extern "C" int crw_labels_upsample(const int32_t* labels, int R, int T, int N, int H, int W, float* out, void* stream) {
    if (!labels || !out || R < 0 || T < 1 || N < 1 || H < 1 || W < 1) return CRW_ERR_INVALID;
    const long long total = (long long)R * H * W;
    if (total == 0) return CRW_OK;
    const long long blocks = (total + 255) / 256;
    lp_upsample_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(labels, R, T, N, H, W, out);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

extern "C" int crw_horizontality_xent(const float* emb, int T, int N, int C, float* xent, void* stream) {
    if (!emb || !xent || T < 1 || N < 1 || C < 2) return CRW_ERR_INVALID;
    if (T == 1) return CRW_OK;
    const size_t smem = (size_t)N * (C + 1) * sizeof(float);
    if (smem > 200 * 1024) return CRW_ERR_UNSUPPORTED;
    CRW_CUDA_RET(cudaFuncSetAttribute(xent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xent_kernel<<<T - 1, 256, smem, (cudaStream_t)stream>>>(emb, T, N, C, xent);
    CRW_LAUNCH_RET();
    return CRW_OK;
}

// ------------------------------------------------------------------------------------------
// Host-streamed EXACT tensor path: `feats` lives in PINNED HOST memory.  Each radargram is cut into up to four segments of frames;
// segment s + 1 is copied into a staging double buffer on the copy stream while segment s runs prep -> filter -> refine.  A query
// frame only needs frame 0 and the ctx frames before it, so a later segment [a, b) is the SUB-SEQUENCE [frame 0 | frames a - ctx .. b):
// its queries sit at sub-sequence frames >= ctx + 1, where window, trimmed candidate numbering and hence W / I are exactly those of
// the whole radargram; the context frames are not emitted (n_min).  Results are bit-identical to crw_labelprop_forward with
// CRW_PREC_TC_EXACT.  The gathers run once over the whole call.
// ------------------------------------------------------------------------------------------
// Segments per radargram: ONE by default -- the pipeline then overlaps the copy of radargram r + 1 with the search of radargram r,
// which is what pays (config 5: 77 MB per radargram against 4.4 ms of search).  Cutting one radargram into segments does not: a
// segment's search lasts one filter item (~65 us at config 3) however short it is, so four segments cost 4 x 110 us against 160 us
// for the whole radargram and the tail after the last copy is as long as before (measured: 0.82 against 0.79 ms for copy-then-search).
// CRW_LP_HOST_SEGS=<n> cuts into up to n segments (used by the tests of the sub-sequence logic).
static void host_exact_segments(int T, int ctx, int& nseg, int& seg) {
    int maxseg = 1;
    { const char* e = getenv("CRW_LP_HOST_SEGS"); if (e && atoi(e) > 0) maxseg = atoi(e); }
    nseg = T / 128;
    if (nseg > maxseg) nseg = maxseg;
    if (nseg < 1) nseg = 1;
    seg = ceil_div(T, nseg);
    if (seg < ctx + 2) { seg = T; nseg = 1; }
    nseg = ceil_div(T, seg);
}
extern "C" size_t crw_labelprop_host_exact_scratch_bytes(int R, int T, int N, int C, int k, int have_topk_out) {
    if (T < 1 || N < 1) return 0;
    int nseg, seg;
    host_exact_segments(T, 20, nseg, seg);                       // (upper bound over ctx: segments only get shorter with larger ctx... sized below with slack)
    const size_t tsub = (size_t)T;                              // a segment never exceeds the radargram
    size_t b = 2 * align_up(tsub * N * C * sizeof(float), 256);                                  // staging double buffer
    b += align_up(crw::lp_x_scratch_bytes(1, (int)tsub, N, C, k, 1), 256);
    if (!have_topk_out) b += 2 * align_up((size_t)R * T * k * N * sizeof(float), 256);
    return b + 512;
}
extern "C" int crw_labelprop_forward_host_exact(const float* feats_host, const float* mask0, int R, int T, int N, int C, int M, int ctx,
                                                float radius, float temp, int k, int mode, int do_normalize, int32_t* labels,
                                                float* masks, float* W_or_null, int32_t* I_or_null, void* scratch,
                                                size_t scratch_bytes, void* stream) {
    if (!feats_host || !mask0 || !labels || !masks) return CRW_ERR_INVALID;
    if (R < 0 || T < 1 || N < 1 || C < 1 || M < 1) return CRW_ERR_INVALID;
    if ((W_or_null == nullptr) != (I_or_null == nullptr)) return CRW_ERR_INVALID;
    if (ctx < 1 || k < 1 || !(radius > 0.0f) || !(temp > 0.0f) || (int64_t)N < k) return CRW_ERR_INVALID;
    if (R == 0) return CRW_OK;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, feats_host) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return CRW_ERR_INVALID;                                   // pageable memory would serialise the copies silently
    }
    const size_t need = crw_labelprop_host_exact_scratch_bytes(R, T, N, C, k, W_or_null != nullptr);
    if (!scratch || scratch_bytes < need) return CRW_ERR_WORKSPACE;
    SideCtx* sc = side_ctx();
    if (!sc) return CRW_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    char* sp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
    const size_t stage_bytes = align_up((size_t)T * N * C * sizeof(float), 256);
    float* stage[2] = {reinterpret_cast<float*>(sp), reinterpret_cast<float*>(sp + stage_bytes)};
    sp += 2 * stage_bytes;
    void* xs = sp;
    sp += align_up(lp_x_scratch_bytes(1, T, N, C, k, 1), 256);
    float* Wt = W_or_null;
    int32_t* It = I_or_null;
    if (!Wt) {
        Wt = reinterpret_cast<float*>(sp);
        sp += align_up((size_t)R * T * k * N * sizeof(float), 256);
        It = reinterpret_cast<int32_t*>(sp);
    }
    GatherParams gp;
    int rc = gather_params(gp, Wt, It, mask0, R, T, N, M, ctx, k, mode, labels, masks);
    if (rc != CRW_OK) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int nseg, seg;
    host_exact_segments(T, ctx, nseg, seg);
    const size_t frame = (size_t)N * C;

    std::lock_guard<std::mutex> lock(sc->mu);
    CRW_CUDA_RET(cudaEventRecord(sc->ev_fork, st));               // earlier work on `st` may still use the scratch
    CRW_CUDA_RET(cudaStreamWaitEvent(sc->s_copy, sc->ev_fork, 0));
    int g = 0;
    for (int rg = 0; rg < R; ++rg) {
        const float* src = feats_host + (size_t)rg * T * frame;
        for (int s = 0; s < nseg; ++s, ++g) {
            const int slot = g & 1;
            const int a = s * seg, b = (a + seg < T) ? a + seg : T;
            if (g >= 2) CRW_CUDA_RET(cudaStreamWaitEvent(sc->s_copy, sc->ev_free[slot], 0));
            int tsub, n_min, out_off;
            if (s == 0) {
                tsub = b; n_min = 1; out_off = 0;
                CRW_CUDA_RET(cudaMemcpyAsync(stage[slot], src, (size_t)b * frame * sizeof(float), cudaMemcpyHostToDevice, sc->s_copy));
            } else {
                tsub = 1 + ctx + (b - a); n_min = ctx + 1; out_off = a - ctx - 1;
                CRW_CUDA_RET(cudaMemcpyAsync(stage[slot], src, frame * sizeof(float), cudaMemcpyHostToDevice, sc->s_copy));
                CRW_CUDA_RET(cudaMemcpyAsync(stage[slot] + frame, src + (size_t)(a - ctx) * frame, (size_t)(b - a + ctx) * frame * sizeof(float),
                                             cudaMemcpyHostToDevice, sc->s_copy));
            }
            CRW_CUDA_RET(cudaEventRecord(sc->ev_copied[slot], sc->s_copy));
            CRW_CUDA_RET(cudaStreamWaitEvent(st, sc->ev_copied[slot], 0));
            if (tsub >= 2) {
                LpXPlanStorage plan;
                const size_t o = ((size_t)rg * T + out_off) * k * N;
                // full-size items (four query tiles each, all sixteen epilogue warps busy) on as many SMs as that takes: a segment's
                // search then lasts one item's time and hides behind the next segment's copy
                int sms_seg = ceil_div(tsub * N, 512);
                if (sms_seg > sms) sms_seg = sms;
                if ((rc = lp_x_prepare(stage[slot], 1, tsub, N, C, ctx, radius, temp, k, do_normalize, Wt + o, It + o, xs, sms_seg, st, plan.bytes,
                                       sizeof(plan.bytes), n_min)) != CRW_OK) return rc;
                if ((rc = lp_x_launch(plan.bytes, 0, lp_x_total_slots(plan.bytes), sms, st)) != CRW_OK) return rc;
            }
            CRW_CUDA_RET(cudaEventRecord(sc->ev_free[slot], st));
        }
    }
    if ((rc = gather_launch_seq(gp, st)) != CRW_OK) return rc;
    return gather_launch_par(gp, st);
}
