/*
 * crw_oracle.c -- plain-C CPU restatement of CRW label propagation.
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never by the product.
 *
 * Follows (reference file:line, all under /root/reference/src):
 *   radius bias            imported/maskedatt.py:232-245, imported/labelprop.py:89-96
 *   logits / trim / top-k  imported/maskedatt.py:157-170
 *   label gather           imported/labelprop.py:82,106-109   (context-trim quirk)
 *   frame loop + argmax    utils.py:134-161
 *   L2 normalise           utils.py:115  (F.normalize, eps 1e-12)
 *   norm     see crw_oracle_l2_normalize (per-lane strided fmaf + xor butterfly)
 *
 * The fp32 operation order is PINNED so that the CUDA fp32 path can be compared
 * bit for bit (the reference itself leaves GEMM summation order, exp and tie
 * order to the library):
 *   dot      warp order: p[l] = fmaf chain over the channels 4l .. 4l+3 (one float4 per "lane", l = 0..31; channels
 *            beyond 128 wrap onto the lanes again), then the xor butterfly 16, 8, 4, 2, 1 of plain adds -- see
 *            crw_oracle_dot.  (One coalesced 16-byte load per lane and five shuffles on a GPU; no staging.)
 *   logit    in band (|j-q| < radius): dot * inv_temp,  inv_temp = 1.0f / temp
 *            (this is what ATen's CUDA `tensor /= python_float` computes);
 *            out of band: (-1e10f) * inv_temp   [dot + -1e10f == -1e10f in fp32]
 *   top-k    descending logit, ties by ascending candidate id (frame-major, node-minor)
 *   softmax  e_j = crw_expf(l_j - l_0); s = ((e_0+e_1)+e_2)+...; w_j = e_j / s
 *   gather   acc = 0; for j = 0..k-1: acc = acc + (label * w_j)   (mul, then add; no fma)
 *   argmax   first maximum
 * Compile with -ffp-contract=off so the compiler never fuses on its own.
 *
 * Parity pin: tests/golden/lp_*.npz (outputs of the live reference) via
 * tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CRW_MASK_BIAS (-1e10f)

/* Pinned exp for x <= 0 (softmax arguments): Cody-Waite reduction + degree-5
 * polynomial (Cephes coefficients), every step a single IEEE fp32 operation. */
float crw_oracle_expf(float x) {
    if (x < -87.0f) return 0.0f;
    float t = x * 1.44269504088896341f;
    float n = rintf(t);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float y = fmaf(p, z, r);
    y = y + 1.0f;
    union { int32_t i; float f; } two_n;
    two_n.i = ((int32_t)n + 127) << 23;
    return y * two_n.f;
}

/* The pinned dot product (see the header): the order one warp computes it in.  "Lane" l (0..31) owns the channels c with
 * (c mod 128) in [4l, 4l+4) -- one float4 of a 128-channel row -- and forms p[l] by sequential fmaf over them in ascending c;
 * the 32 partials are then combined by the xor butterfly of the normalisation (offsets 16, 8, 4, 2, 1; plain adds), whose
 * result is the same on every lane because fp addition is commutative. */
float crw_oracle_dot(const float* a, const float* b, int C) {
    float p[32], t[32];
    for (int l = 0; l < 32; ++l) {
        float acc = 0.0f;
        for (int base = 0; base < C; base += 128)
            for (int i = 0; i < 4; ++i) {
                const int c = base + 4 * l + i;
                if (c < C) acc = fmaf(a[c], b[c], acc);
            }
        p[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; ++l) p[l] = t[l];
    }
    return p[0];
}

int crw_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* x [rows, C] -> out [rows, C] = x / max(||x||_2, 1e-12).  Pinned order (maps onto one warp per
 * row): partial[l] = sequential fmaf over c = l, l+32, l+64, ...; then a 5-stage xor butterfly
 * (offsets 16,8,4,2,1) of plain adds; nrm = sqrtf(total); out = x / max(nrm, eps). */
void crw_oracle_l2_normalize(const float* x, int64_t rows, int C, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) {
        const float* xi = x + i * C;
        float p[32], t[32];
        for (int l = 0; l < 32; ++l) {
            float ss = 0.0f;
            for (int c = l; c < C; c += 32) ss = fmaf(xi[c], xi[c], ss);
            p[l] = ss;
        }
        for (int off = 16; off >= 1; off >>= 1) {
            for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ off];
            for (int l = 0; l < 32; ++l) p[l] = t[l];
        }
        float d = fmaxf(sqrtf(p[0]), 1e-12f);
        for (int c = 0; c < C; ++c) out[i * C + c] = xi[c] / d;
    }
}

static inline int n_key_frames(int n, int ctx) { return n <= ctx + 1 ? n : ctx + 1; }
/* f-th key frame of query frame n (trim of maskedatt.py:166-167). */
static inline int key_frame(int n, int ctx, int f) {
    if (n <= ctx + 1) return f;
    return f == 0 ? 0 : n - ctx + (f - 1);
}
/* frame whose soft mask candidate-frame slot f is gathered from (labelprop.py:82,106). */
static inline int label_frame(int n, int ctx, int f, int mode_fixed) {
    if (n <= ctx + 1 || mode_fixed) return key_frame(n, ctx, f);
    return f; /* quirk: untrimmed list, first (ctx+1) frames */
}

/* Top-k + softmax for query frames n in [n_begin, n_end) of one radargram.
 * emb [T,N,C] (normalised), W [T,k,N] f32, I [T,k,N] i32 (rows of frame 0 untouched). */
int crw_oracle_lp_topk(const float* emb, int T, int N, int C, int ctx, float radius, float temp,
                       int k, int n_begin, int n_end, float* W, int32_t* I) {
    if (T < 1 || N < 1 || C < 1 || ctx < 1 || k < 1 || k > 64) return -1;
    if (n_begin < 1) n_begin = 1;
    if (n_end > T) n_end = T;
    const float inv_temp = 1.0f / temp;
    const float masked = CRW_MASK_BIAS * inv_temp;
#pragma omp parallel for schedule(dynamic, 4)
    for (int n = n_begin; n < n_end; ++n) {
        const int F = n_key_frames(n, ctx);
        if (F * N < k) continue; /* torch.topk would raise; caller validates */
        float bv[64];
        int32_t bi[64];
        for (int q = 0; q < N; ++q) {
            const float* qv = emb + ((int64_t)n * N + q) * C;
            int cnt = 0;
            for (int f = 0; f < F; ++f) {
                const float* kf = emb + (int64_t)key_frame(n, ctx, f) * N * C;
                for (int j = 0; j < N; ++j) {
                    float logit;
                    int dj = j - q; if (dj < 0) dj = -dj;
                    if ((float)dj < radius) {
                        const float* kv = kf + (int64_t)j * C;
                        logit = crw_oracle_dot(kv, qv, C) * inv_temp;
                    } else {
                        logit = masked;
                    }
                    /* insertion keeping (logit desc, id asc): a later id never displaces an equal logit */
                    int pos = cnt;
                    if (cnt == k) { if (!(logit > bv[k - 1])) continue; pos = k - 1; }
                    else cnt++;
                    while (pos > 0 && logit > bv[pos - 1]) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                    bv[pos] = logit; bi[pos] = f * N + j;
                }
            }
            float e[64], s = 0.0f;
            for (int j = 0; j < k; ++j) { e[j] = crw_oracle_expf(bv[j] - bv[0]); s = (j == 0) ? e[0] : s + e[j]; }
            for (int j = 0; j < k; ++j) {
                W[((int64_t)n * k + j) * N + q] = e[j] / s;
                I[((int64_t)n * k + j) * N + q] = bi[j];
            }
        }
    }
    return 0;
}

/* Label gather over all frames of one radargram.
 * mask0 [M,N]; masks [T,M,N] out; labels [T,N] out (i32); mode_fixed: 0 = ref_exact quirk. */
int crw_oracle_lp_gather(const float* W, const int32_t* I, const float* mask0, const int32_t* label0,
                         int T, int N, int M, int ctx, int k, int mode_fixed, float* masks,
                         int32_t* labels) {
    if (T < 1 || M < 1 || M > 64) return -1;
    memcpy(masks, mask0, sizeof(float) * (size_t)M * N);
    for (int q = 0; q < N; ++q) labels[q] = label0[q];
    for (int n = 1; n < T; ++n) {
        for (int q = 0; q < N; ++q) {
            float acc[64];
            for (int m = 0; m < M; ++m) acc[m] = 0.0f;
            for (int j = 0; j < k; ++j) {
                const int32_t id = I[((int64_t)n * k + j) * N + q];
                const float w = W[((int64_t)n * k + j) * N + q];
                const int lf = label_frame(n, ctx, id / N, mode_fixed);
                const float* src = masks + (int64_t)lf * M * N + (id % N);
                for (int m = 0; m < M; ++m) {
                    float prod = src[(int64_t)m * N] * w;
                    acc[m] = acc[m] + prod;
                }
            }
            int best = 0;
            for (int m = 0; m < M; ++m) {
                masks[((int64_t)n * M + m) * N + q] = acc[m];
                if (acc[m] > acc[best]) best = m;
            }
            labels[(int64_t)n * N + q] = best;
        }
    }
    return 0;
}

/* Whole path for R radargrams: feats [R,T,N,C] raw encoder output (normalised here
 * when do_normalize), mask0 [R,M,N], label0 [R,N] -> labels [R,T,N], masks [R,T,M,N] (or NULL). */
int crw_oracle_labelprop(const float* feats, const float* mask0, const int32_t* label0, int R, int T,
                         int N, int C, int M, int ctx, float radius, float temp, int k, int mode_fixed,
                         int do_normalize, int32_t* labels, float* masks_or_null,
                         float* W_or_null, int32_t* I_or_null) {
    const int64_t per = (int64_t)T * N * C;
    float* emb = (float*)malloc(sizeof(float) * per);
    float* W = W_or_null ? NULL : (float*)malloc(sizeof(float) * (size_t)T * k * N);
    int32_t* I = I_or_null ? NULL : (int32_t*)malloc(sizeof(int32_t) * (size_t)T * k * N);
    float* masks = masks_or_null ? NULL : (float*)malloc(sizeof(float) * (size_t)T * M * N);
    int rc = 0;
    for (int r = 0; r < R && rc == 0; ++r) {
        float* Wr = W_or_null ? W_or_null + (int64_t)r * T * k * N : W;
        int32_t* Ir = I_or_null ? I_or_null + (int64_t)r * T * k * N : I;
        float* mr = masks_or_null ? masks_or_null + (int64_t)r * T * M * N : masks;
        if (do_normalize) crw_oracle_l2_normalize(feats + r * per, (int64_t)T * N, C, emb);
        else memcpy(emb, feats + r * per, sizeof(float) * per);
        memset(Wr, 0, sizeof(float) * (size_t)k * N);
        memset(Ir, 0, sizeof(int32_t) * (size_t)k * N);
        rc = crw_oracle_lp_topk(emb, T, N, C, ctx, radius, temp, k, 1, T, Wr, Ir);
        if (rc == 0)
            rc = crw_oracle_lp_gather(Wr, Ir, mask0 + (int64_t)r * M * N, label0 + (int64_t)r * N, T, N, M,
                                      ctx, k, mode_fixed, mr, labels + (int64_t)r * T * N);
    }
    free(emb); free(W); free(I); free(masks);
    return rc;
}
