/*
 * crw_b200.h -- C ABI of libcrw_b200.so: the B200 (sm_100a) CRW hot path.
 *
 * The reference (jdalcorso/radar-sounder-crw) has no FFI of its own: its hot path is
 * Python calling ATen.  The entry points below are what a maintainer would bind from
 * the reference's call sites (ctypes stubs are shown in INTEGRATION.md); each one names
 * the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer (cudaMalloc'd / torch CUDA tensor
 *     storage); the library never allocates or frees device memory for tensors and
 *     never synchronises with the host (workspaces come from the caller, sized by the
 *     crw_*_bytes functions);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) and is ordered after / before other work on that stream;
 *   - process-wide state the library DOES keep, per device, created on first use and
 *     never freed: a higher-priority side stream, a copy stream and a handful of
 *     events (crw_labelprop_forward forks the early frames + the sequential label
 *     gather onto the side stream and joins them back with events; the enqueue
 *     sequence of a call is serialised by a mutex, so calls from several host threads
 *     are safe but do not interleave) and cached kernel attributes (opt-in shared-memory
 *     sizes).  The environment switches listed in DESIGN.md (engine selection for tests,
 *     profiling aids) are looked up with getenv at every call -- none is needed in
 *     production and none changes results except the documented debug flags;
 *   - return value: CRW_OK, a negative CRW_ERR_* code, or -(1000 + cudaError_t) when
 *     a launch failed; no exceptions cross the boundary;
 *   - tensors are dense row-major fp32 unless said otherwise; shapes in brackets.
 *
 * Symbols:  B batch, T frames, N nodes per frame, C channels, M classes, R radargrams,
 *           k top-k, ctx context frames, tau / temp temperature.
 */
#ifndef CRW_B200_H_
#define CRW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRW_OK 0
#define CRW_ERR_INVALID (-1)     /* bad shape / parameter / null pointer            */
#define CRW_ERR_UNSUPPORTED (-2) /* valid request this build cannot serve (k > 32 ...) */
#define CRW_ERR_ALIGN (-3)       /* pointer not 16-byte aligned or C % 4 != 0         */
#define CRW_ERR_WORKSPACE (-4)   /* workspace too small                               */
#define CRW_ERR_CUDA_BASE (-1000)

/* precision selectors */
#define CRW_PREC_FP32 0        /* fp32 FMA, pinned order: bit-comparable with oracle/crw_oracle.c */
#define CRW_PREC_BF16X3 1      /* tcgen05 kind::f16, error-compensated bf16 hi/lo (3 MMAs), fp32 accumulate */
/* (value 2 is unassigned: a kind::tf32 walk was declared in round 1 and never built; error-compensated bf16 covers that case) */
#define CRW_PREC_TC_EXACT 3    /* label propagation only: one tcgen05 fp16 pass FILTERS the candidates with a proven margin, the
                                  survivors are re-scored in fp32 in the pinned order -- results bit-identical to CRW_PREC_FP32 */

/* label-propagation gather modes */
#define CRW_LP_REF_EXACT 0     /* reproduce the reference's context-trim gather quirk */
#define CRW_LP_FIXED 1         /* gather from the frames the keys came from           */

int crw_version(void);
const char* crw_error_string(int code);
/* compute capability the library was built for (100 = sm_100a) */
int crw_built_arch(void);

/* ------------------------------------------------------------------------------------
 * L2 normalisation -- replaces F.normalize(emb, dim=-1): src/model.py:22, src/utils.py:115
 *   x, out [rows, C]; out may alias x.
 * ---------------------------------------------------------------------------------- */
int crw_l2_normalize(const float* x, int64_t rows, int C, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Training walk -- replaces the tail of CRW.forward, src/model.py:22-46:
 *   normalise, stride-1 affinities / tau (model.py:26), palindrome walk (model.py:31-44),
 *   cycle cross-entropy (model.py:45), return loss/N (model.py:46).
 *   x        [B,T,N,C]  raw encoder output (un-normalised)
 *   loss     [1]        loss/N
 *   A        [B,T-1,N,N] or NULL   the affinities the reference also returns (model.py:46)
 *   saved    workspace of crw_walk_saved_bytes(): state kept for the backward pass
 * T < 3 gives loss = 0 exactly as the reference's empty loop (model.py:33-35).
 * ---------------------------------------------------------------------------------- */
size_t crw_walk_saved_bytes(int B, int T, int N, int C, int precision);
int crw_walk_forward(const float* x, int B, int T, int N, int C, float tau, int precision,
                     float* loss, float* A_or_null, void* saved, size_t saved_bytes, void* stream);

/* Reverse pass -- replaces autograd through src/model.py:22-46 (scripts/train.py:71).
 *   x        [B,T,N,C] the same raw encoder output given to crw_walk_forward
 *   saved    the workspace crw_walk_forward filled
 *   dloss    [1] device scalar: upstream gradient of loss
 *   dA       [B,T-1,N,N] or NULL: upstream gradient of the returned affinities
 *   dx       [B,T,N,C] out: gradient w.r.t. the raw encoder output
 *   scratch  workspace of crw_walk_backward_scratch_bytes()
 * ---------------------------------------------------------------------------------- */
size_t crw_walk_backward_scratch_bytes(int B, int T, int N, int C, int precision);
int crw_walk_backward(const float* x, const void* saved, size_t saved_bytes, const float* dloss,
                      const float* dA_or_null, int B, int T, int N, int C, float tau, int precision, float* dx,
                      void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Affinity + top-k + softmax -- replaces batched_affinity, src/imported/maskedatt.py:151-175,
 * with the radius bias of maskedatt.py:232-245 / labelprop.py:89-96 applied to every key
 * frame and the context trim of maskedatt.py:166-167 applied *before* any arithmetic.
 *   keys     [n_keys,N,C]  normalised features of ALL previous frames (frame 0 first)
 *   queries  [n_q,N,C]     normalised features of query frames; query i plays frame
 *                          n = n_first + i and sees keys 0..n-1 (trimmed to frame 0 + last ctx)
 *   W        [n_q,k,N] f32 softmax weights, sorted by descending logit
 *   I        [n_q,k,N] i32 candidate ids into the trimmed key set (frame-slot * N + node)
 * Ties: equal logits ordered by ascending id (torch.topk leaves this unspecified).
 * Whole-sequence use: keys = emb, queries = emb + N*C, n_first = 1, n_q = T-1.
 * ---------------------------------------------------------------------------------- */
int crw_affinity_topk(const float* keys, const float* queries, int n_first, int n_q, int N, int C,
                      int ctx, float radius, float temp, int k, int precision, float* W, int32_t* I,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * Label gather + argmax -- replaces the tail of LabelPropVOS_CRW.predict
 * (src/imported/labelprop.py:82,106-116) and the frame loop / argmax of propagate
 * (src/utils.py:152-160) for R independent radargrams.
 *   W, I     [R,T,k,N]  as written by crw_affinity_topk with n_first = 1 (row 0 unused)
 *   mask0    [R,M,N]    one-hot reference mask of frame 0 (utils.py:143-147)
 *   labels   [R,T,N] i32 out: argmax class per node (labels[:,0] = argmax of mask0)
 *   masks    [R,T,M,N]  out: soft masks of every frame (the reference keeps them, utils.py:157)
 * ---------------------------------------------------------------------------------- */
int crw_label_gather(const float* W, const int32_t* I, const float* mask0, int R, int T, int N, int M,
                     int ctx, int k, int mode, int32_t* labels, float* masks, void* stream);

/* One stepwise predict call -- replaces src/imported/labelprop.py:106-116 for callers that keep
 * the reference's per-frame API (LabelPropVOS_CRW.predict):
 *   W, I [k,N] from crw_affinity_topk (n_q = 1); lbl [F,M,N] soft masks of the frames the ids
 *   index into (frame-slot major); out_mask [M,N]; out_label [N] i32 or NULL.
 */
int crw_label_gather_step(const float* W, const int32_t* I, const float* lbl, int F, int N, int M, int k,
                          float* out_mask, int32_t* out_label_or_null, void* stream);

/* ------------------------------------------------------------------------------------
 * Whole label-propagation path for R radargrams -- replaces propagate's encode-side tail and
 * frame loop, src/utils.py:115,134-161 (normalise, top-k, gather, argmax).
 *   feats    [R,T,N,C]  raw encoder output (normalised here when do_normalize != 0)
 *   mask0    [R,M,N]
 *   labels   [R,T,N] i32 out
 *   masks    [R,T,M,N] out (required: the recurrence reads it)
 *   W_or_null, I_or_null [R,T,k,N]: top-k dumps; when NULL they live in `scratch`
 *   scratch  workspace of crw_labelprop_scratch_bytes()
 * ---------------------------------------------------------------------------------- */
size_t crw_labelprop_scratch_bytes(int R, int T, int N, int C, int k, int precision, int do_normalize,
                                   int have_topk_out);
int crw_labelprop_forward(const float* feats, const float* mask0, int R, int T, int N, int C, int M,
                          int ctx, float radius, float temp, int k, int mode, int precision,
                          int do_normalize, int32_t* labels, float* masks, float* W_or_null,
                          int32_t* I_or_null, void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Same as crw_labelprop_forward (tensor path, precision = CRW_PREC_BF16X3) with the features in PINNED HOST memory:
 * every radargram is streamed to the device in chunks of whole query tiles on an internal copy stream while the previous
 * chunk is normalised, split and searched on `stream` (a query tile only needs key rows that precede it).  This is the
 * call for a caller whose encoder output / feature cache lives on the host; the host buffer must stay untouched until
 * the work queued on `stream` has completed.  Returns CRW_ERR_INVALID for pageable memory.
 *   scratch: crw_labelprop_host_scratch_bytes() bytes (adds a staging double buffer)
 * ---------------------------------------------------------------------------------- */
size_t crw_labelprop_host_scratch_bytes(int R, int T, int N, int C, int k, int have_topk_out);
int crw_labelprop_forward_host(const float* feats_host, const float* mask0, int R, int T, int N, int C, int M,
                               int ctx, float radius, float temp, int k, int mode, int do_normalize,
                               int32_t* labels, float* masks, float* W_or_null, int32_t* I_or_null, void* scratch,
                               size_t scratch_bytes, void* stream);
/* The same for the EXACT tensor path (CRW_PREC_TC_EXACT): each radargram is cut into up to four segments of frames, segment s + 1
 * is copied while segment s runs prep -> filter -> refine; a later segment is the sub-sequence [frame 0 | its ctx context frames |
 * its frames], whose windows and candidate numbering are those of the whole radargram -- results bit-identical to
 * crw_labelprop_forward(..., CRW_PREC_TC_EXACT, ...). */
size_t crw_labelprop_host_exact_scratch_bytes(int R, int T, int N, int C, int k, int have_topk_out);
int crw_labelprop_forward_host_exact(const float* feats_host, const float* mask0, int R, int T, int N, int C, int M,
                               int ctx, float radius, float temp, int k, int mode, int do_normalize,
                               int32_t* labels, float* masks, float* W_or_null, int32_t* I_or_null, void* scratch,
                               size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * "Horizontality" metric -- replaces src/utils.py:118-123 (channel-shifted intra-frame
 * similarity / 0.1, cross-entropy against the identity, reduction='none').
 *   emb [T,N,C] normalised  ->  xent [N,T-1]
 * ---------------------------------------------------------------------------------- */
int crw_horizontality_xent(const float* emb, int T, int N, int C, float* xent, void* stream);

/* ------------------------------------------------------------------------------------
 * Nearest-neighbour upsample of the propagated label map to pixel resolution -- replaces
 * `up = Resize((seg_h, rg_len), NEAREST)` applied to final_prediction, scripts/test/test_all.py:79,96.
 *   labels [R,T,N] i32  ->  out [R,H,W] f32  (out[r,y,x] = labels[r, floor(x*T/W), floor(y*N/H)])
 * ---------------------------------------------------------------------------------- */
int crw_labels_upsample(const int32_t* labels, int R, int T, int N, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Radargram -> frame sequences -- replaces RGDataset.__getitem__ / get_smaller_item, src/dataset.py:34-47
 * (slice, unfold rows, unfold columns, permute) and the `use_last` frame flip of src/utils.py:108.
 *   rg  [H, ld] f32, device resident (the `.pt` radargram of src/utils.py:32-38 copied once)
 *   item r starts at column col_start + r*col_stride (dataset index i: col_start = (w-ow)*i, dataset.py:35)
 *   out [R,T,N,h,w]:  out[r,t,n,y,x] = rg[n*(h-oh)+y][col_start + r*col_stride + t'*(w-ow) + x],
 *                     t' = T-1-t when `reverse`, else t.   N*h - oh*(N-1) <= H is required (dataset.py:27).
 * ---------------------------------------------------------------------------------- */
int crw_patch_unfold(const float* rg, int H, int64_t ld, int64_t col_start, int64_t col_stride, int R, int T, int N,
                     int h, int w, int oh, int ow, int reverse, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * First-column label seeding -- replaces `down = Resize((N,1), NEAREST)`, the label read and the
 * one-hot mask loop of src/utils.py:139-147 (seg_ref = seg[:rows, col : col+W], scripts/test/test_all.py:94).
 *   seg [rows.., ld] f32 class ids; radargram r reads column col_start + r*col_stride
 *   label0_or_null [R,N] i32:  seg[min(floor(i * float(rows)/float(N)), rows-1)][col]
 *   mask0_or_null  [R,M,N] f32 one-hot (the `mask0` input of crw_labelprop_forward)
 * ---------------------------------------------------------------------------------- */
int crw_seed_labels(const float* seg, int rows, int64_t ld, int64_t col_start, int64_t col_stride, int R, int N, int M,
                    int32_t* label0_or_null, float* mask0_or_null, void* stream);

/* ------------------------------------------------------------------------------------
 * Reversed-pass mask fusion -- replaces scripts/test/test_all.py:146-158: the reversed pass's map is un-flipped per
 * radargram (unfold / flip / view) and its class-2 (bedrock) pixels overwrite the forward map.
 *   fwd, rev, out [H, W] f32 class ids, W a multiple of rg_len; `rev` in the reversed pass's own column order;
 *   out may alias fwd.   rule = the reference's dataset id:
 *     0: mask = rev' == 2                         1: ... && fwd != 3 && no class 4 anywhere in that column of rev'
 *     3: mask = rev' == 2, second half of the flattened map only
 *   scratch: crw_fuse_reversed_scratch_bytes(W) bytes (used by rule 1).
 * ---------------------------------------------------------------------------------- */
size_t crw_fuse_reversed_scratch_bytes(int64_t W);
int crw_fuse_reversed(const float* fwd, const float* rev, int H, int64_t W, int rg_len, int rule, float* out,
                      void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimizer step of the train loop -- replaces `optimizer.step()` of scripts/train.py:56,72 (torch.optim.Adam(lr), amsgrad
 * and maximize off) for ALL parameters at once: params / grads / exp_avg / exp_avg_sq are flat f32 device buffers of n
 * elements (16-byte aligned; radar_sounder_crw_b200.optim.FlatAdam lays every parameter and its gradient out as views of
 * such buffers), one elementwise launch.  `step` = 1 for the first call (bias corrections 1 - beta^step are computed in
 * double on the host from the double hyper-parameters, as torch does with its python floats); the gradient is read as grad_scale * grads (1 / world after a SUM all-reduce).
 *   g' = grad_scale g + weight_decay p;  m += (g' - m)(1 - beta1);  v = beta2 v + (1 - beta2) g'^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * ---------------------------------------------------------------------------------- */
int crw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int64_t step, double grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Self-test of the tensor-core plumbing (TMA SWIZZLE_128B tiles -> tcgen05.mma -> TMEM -> tcgen05.ld):
 *   out[128,BN] = A[128,128] * B[BN,128]^T, A/B bf16 row-major, out fp32; BN multiple of 16, <= 256.
 * No reference counterpart; it pins the descriptor encodings the tensor-core kernels rely on.
 * ---------------------------------------------------------------------------------- */
int crw_debug_umma_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream);
/* same product with A parked in TMEM by the threads (tcgen05.st) and read by the "TS" form of tcgen05.mma */
int crw_debug_umma_ts_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream);
/* same product with A brought in by TMA and copied shared memory -> TMEM by tcgen05.cp (128x256b per K = 16 step), then read
 * by TS MMAs; BN multiple of 32, 32..256.  Pins the smem -> TMEM copy against the A-in-TMEM layout. */
int crw_debug_umma_tscp_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream);
/* K = 64 product with MN-major operands: A is At[64][128] when a_mn else A[128][64]; B is Bkn[64][BN] when b_mn else
 * Bt[BN][64]; BN in {64, 128}.  Pins the MN-major shared-memory descriptors (transposed operands without a copy). */
int crw_debug_umma_mn_gemm(const void* A_bf16, const void* B_bf16, int BN, int a_mn, int b_mn, float* out, void* stream);
/* CTA-pair form (tcgen05 cta_group::2 on a 2-CTA cluster): out[256,BN] = A[256,128] * B[BN,128]^T, BN multiple of 32,
 * 32..256.  Pins the pair plumbing (leader-credited TMA, M=256 MMA, multicast commit) of the label-propagation kernel. */
int crw_debug_umma_pair_gemm(const void* A_bf16, const void* B_bf16, int BN, float* out, void* stream);
/* Profiling aid for the tensor-path top-k kernel: with env CRW_TC_DEBUG bit 3 set the epilogue warps accumulate cycles per
 * phase; this copies the 160 x 8 x 10 uint64 counters to HOST memory (synchronising) and clears them when `reset`. */
int crw_debug_lp_profile(unsigned long long* host_out, int reset);
/* same for the exact tensor path (CRW_PREC_TC_EXACT): [2 kernels: filter, refine][160 CTAs][18 warps][8 phases] uint64 */
int crw_debug_lp_x_profile(unsigned long long* host_out, int reset);
/* profiling aid: per-phase cycle counters of the fused walk kernels, [forward, backward][16] uint64 (CRW_WALK_PROF=1) */
int crw_debug_walk_fused_profile(unsigned long long* host_out, int reset);
/* Test aid, host only (no GPU): the work items of a tensor-path top-k launch over n_tiles query tiles on `grid` CTAs -- whole tiles,
 * and, when the last round is partial, its tiles cut in two key-range halves scheduled first.  out[3 i .. 3 i + 2] = tile, half
 * (-1 whole, 0 / 1), index among the split tiles; returns the number of items (out may be null to query it), < 0 on error. */
int crw_debug_lp_schedule(int n_tiles, int grid, int split_ok, int* out, int out_capacity_items);

#ifdef __cplusplus
}
#endif
#endif /* CRW_B200_H_ */
