/*
 * vlgae_b200.h -- C ABI of libvlgae_b200.so (sm_100a).
 *
 * The drop-in boundary for VLGAE's structured-inference hot path.  The
 * reference (LouChao98/VLGAE) is pure Python: its "FFI" for this path is the
 * operator API of src/model/torch_struct and the gather_logit implementation
 * group of src/model/joint.py.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root).  The Python
 * mirror of that API (vlgae_b200/torch_struct, vlgae_b200/alignment.py) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes; every data pointer is a DEVICE pointer unless
 *     the name ends in _host; tensors are contiguous, row-major, float32
 *     (lengths: int64) exactly as the reference holds them.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     Calls only enqueue work; nothing synchronises the host.
 *   - return 0 on success; nonzero = VLGAE_E_*; vlgae_last_error() gives the
 *     text (thread-local).  There is no CPU fallback anywhere.
 *
 * Tensor layouts (reference: src/model/torch_struct/dmv.py:24-31)
 *   dec     [B][N][2 dir][2 val][2 decision]   merged (ROOT at position 0)
 *   attach  [B][N][N][2 val]   (head, child, valence)   merged
 *   lengths [B]  int64, words without ROOT, 0 <= len <= N-1
 *   constants: LEFT=0 RIGHT=1, HASCHILD=0 NOCHILD=1, GO=0 STOP=1 (dmv.py:7-15)
 */
#ifndef VLGAE_B200_H_
#define VLGAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLGAE_OK 0
#define VLGAE_E_INVALID 1   /* bad argument (null pointer, N out of range, ...) */
#define VLGAE_E_CUDA 2      /* a CUDA runtime call failed                        */
#define VLGAE_E_WORKSPACE 3 /* workspace too small                               */
#define VLGAE_E_ARCH 4      /* device is not sm_100                              */

#define VLGAE_DMV_MAX_N 256 /* chart positions incl. ROOT (reference data: <= 51) */
#define VLGAE_ALIGN_MAX_D 128 /* feature width of the alignment contraction (reference: 128) */

/* ABI version (bumped on any signature change). */
int vlgae_version(void);
/* Text of the last error on this thread. */
const char *vlgae_last_error(void);

/*
 * Schedule of the DMV kernels (process-wide): 0 = automatic, 1 = frontier (one thread per target cell, running
 * log-sum-exp / arg-max state; the latency regime, short sentences, charts beyond 41 positions, host-memory hand-off),
 * 2 = gather (lanes stream the split points of a span; log semiring in the linear domain on per-word offset scores with
 * a self-check, value-only Viterbi with a re-evaluating back-trace; csrc/dmv_gather.cu).  Automatic = gather for
 * batches beyond one resident wave padded to 28 .. 41 positions, frontier otherwise.  A schedule that cannot run a
 * launch (chart beyond its layout, no workspace for the redo flags) falls back to the automatic choice.  Results agree
 * within the documented tolerances (max semiring: bit-exact).  Used by the tests and the sweep tool.
 */
int vlgae_dmv_set_schedule(int which);

/*
 * Frontier schedule: sentences of at most `words` words run their log-semiring sweeps in the LINEAR domain (sums of
 * products on exp(offset scores), self-checked, log-domain fallback inside the CTA; csrc/dmv_frontier.cu, LIN).
 * words < 0 restores the default (24, or VLGAE_FRONTIER_LINEAR); 0 = never; a large value = every length.  Unless 0,
 * sentences of >= 45 words (charts of 46..72 positions) also take the linear sweeps: there the reference itself is
 * > 1e-5 from the exact result and parity is judged three-way (GPU error <= reference error) in every test.
 * The linear sweeps are faster (cfg2 batch 53 -> 45 us) and closer to the exact marginals (2.5e-7 instead of 1.1e-6
 * from fp64); the default is length-bound because the reference's own fp32 sweep (torch_struct/dmv.py:47-63 +
 * helpers.py:150-154) drifts to 1.0e-5 from the exact result at 40 words, and a result that is closer to the truth
 * than that cannot also stay within 1e-5 of the reference.  Process-wide, like vlgae_dmv_set_schedule.
 */
int vlgae_dmv_set_linear_max_len(int words);

/*
 * Debug aid: a device buffer of 8 int64; the CTAs of sentence 0 write cumulative SM cycle counts after each phase
 * ([0..3] log pass: staged, inside done, outside done, outputs written; [4..6] max pass: staged, chart done,
 * back-trace done).  NULL disables it (default).
 */
int vlgae_dmv_set_profile_buffer(void *device_buf);

/*
 * Bytes of device scratch the DMV entry points need for a batch of B sentences
 * padded to N positions: B redo flags of the gather schedule (a sentence whose
 * linear-domain sweep fails its self-check is flagged and redone in the log domain
 * by a follow-up launch), plus, when the chart does not fit in shared memory
 * (N > ~72), one chart slice per resident CTA (L2-resident).  Without a workspace
 * the entry points keep to the frontier schedule.
 */
size_t vlgae_dmv_workspace_bytes(int B, int N);

/*
 * Log semiring: Z and the expected counts.
 * Replaces  DMV1o(...).partition              src/model/torch_struct/distributions.py:190-193
 *           -> _Struct.sum -> DMV1oStruct._dp  helpers.py:101-116, dmv.py:19-66
 *     and   torch.autograd.grad(partition.sum(), [dec, attach]) / DMV1o(...).marginals
 *           src/model/torch_struct/helpers.py:118-154 (autograd through the chart),
 *           here an explicit reverse sweep in the same kernel.
 *   gZ      [B] upstream gradient of Z (NULL = ones)
 *   Z       [B]                       (the reference returns [B,1]; same memory)
 *   gdec    [B][N][2][2][2] or NULL   d(sum_b gZ[b] Z[b]) / d dec
 *   gattach [B][N][N][2]    or NULL   d(...) / d attach  (= arc marginals when gZ = 1)
 * Padded positions receive exact zeros.  If both gdec and gattach are NULL only
 * the inside pass runs.
 * mask_zero: value written by the single-root mask (dmv.py:63; class attr zero = -1e12).
 */
int vlgae_dmv_inside_outside(const float *dec, const float *attach, const int64_t *lengths, int B, int N,
                             float mask_zero, const float *gZ, float *Z, float *gdec, float *gattach, void *workspace,
                             size_t workspace_bytes, void *stream);

/*
 * Max semiring: best score, heads, dense arc indicator, decision counts.
 * Replaces  DMV1o(...).max     distributions.py:116-123
 *           DMV1o(...).argmax  distributions.py:125-133 -> helpers.py:118-154 with MaxSemiring
 *                              (semirings.py:187-207; torch.max first-index tie rule)
 *   best   [B]
 *   heads  [B][N] int64 or NULL: heads[b][c] = head of word c (1..len), 0 elsewhere
 *                                (what callers build from argmax.sum(-1).nonzero(),
 *                                 ldndmv.py:301-303, joint.py:256-258)
 *   arcs   [B][N][N][2] or NULL: the dense 0/1 tensor `argmax` returns (= d max / d attach)
 *   gdec   [B][N][2][2][2] or NULL: d max / d dec (decision counts of the best tree)
 */
int vlgae_dmv_viterbi(const float *dec, const float *attach, const int64_t *lengths, int B, int N, float mask_zero,
                      float *best, int64_t *heads, float *arcs, float *gdec, void *workspace, size_t workspace_bytes,
                      void *stream);

/*
 * Both of the above in ONE launch (log-semiring CTAs and max-semiring CTAs run side by side):
 * the "inside + outside + Viterbi" step BASELINE.json's metric is quoted on
 * (src/model/joint.py:251-256 runs exactly this pair per training step).
 */
int vlgae_dmv_parse(const float *dec, const float *attach, const int64_t *lengths, int B, int N, float mask_zero,
                    const float *gZ, float *Z, float *gdec, float *gattach, float *best, int64_t *heads, float *arcs,
                    float *vgdec, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Same as vlgae_dmv_parse but with HOST buffers; synchronises `stream` before returning (the
 * end-to-end call a non-torch embedder would make).  Any output pointer may be NULL.
 * Pinned buffers (cudaHostAlloc / cudaHostRegister, torch pin_memory()) are read and written by
 * the kernel directly over PCIe -- no staging copies; pageable buffers are staged through a
 * grow-only device arena owned by the calling host thread.  VLGAE_ZERO_COPY=0 forces staging.
 */
int vlgae_dmv_parse_host(const float *dec_host, const float *attach_host, const int64_t *lengths_host, int B, int N,
                         float mask_zero, float *Z_host, float *gdec_host, float *gattach_host, float *best_host,
                         int64_t *heads_host, void *stream);

/*
 * The same call without the final synchronisation: it returns once the work is enqueued on `stream`
 * (the stream semantics every torch operator of the reference has -- torch_struct/dmv.py runs
 * asynchronously on the current stream and the caller synchronises when it reads a result).  The
 * host buffers must stay valid and untouched until `stream` has drained; the results are in host
 * memory from then on.  Calls on different streams may be in flight together (bulk decoding with
 * two batches double-buffered: the PCIe traffic of one overlaps the sweeps of the other).
 * Zero-copy only: every buffer must be pinned, Z_host and best_host non-null and N within the
 * shared-memory charts (N <= 72), otherwise VLGAE_E_INVALID; B > 256 is fine.
 */
int vlgae_dmv_parse_host_async(const float *dec_host, const float *attach_host, const int64_t *lengths_host, int B, int N,
                               float mask_zero, float *Z_host, float *gdec_host, float *gattach_host, float *best_host,
                               int64_t *heads_host, void *stream);

/*
 * DMV1o.merge  (distributions.py:253-265): prepend ROOT.
 *   dec [B][n][2][2][2], attach [B][n][n][2], root [B][n]  ->  dec_w [B][n+1][2][2][2], attach_w [B][n+1][n+1][2]
 */
int vlgae_dmv_merge(const float *dec, const float *attach, const float *root, int B, int n, float one, float zero,
                    float *dec_w, float *attach_w, void *stream);

/*
 * Arc-factored projective dependency CRF (MBR decoding).
 * Replaces  DependencyCRF(arc, lengths).partition / .max / .marginals / .argmax
 *           src/model/torch_struct/distributions.py:269-299 -> deptree.py:25-76,146-162 (+ helpers.py:118-154)
 *   arc [B][N][N] (head, child); positions beyond lengths[b] are treated as `fill` (deptree.py:159-161; the
 *   reference's run-time global NEGINF); mask_zero = value of the single-root mask (class attr zero, deptree.py:72-73).
 *   semiring 0 = log: out[b] = log Z, marginals = arc marginals;  1 = max: out[b] = best score, marginals = 0/1
 *   indicator of the best tree (torch.max first-index tie rule), heads[b][c] = head of word c (0 elsewhere).
 *   marginals / heads may be NULL (then only the forward sweep runs).
 */
size_t vlgae_deptree_workspace_bytes(int B, int N);
int vlgae_deptree(const float *arc, const int64_t *lengths, int B, int N, float fill, float mask_zero, int semiring,
                  float *out, float *marginals, int64_t *heads, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Alignment scores (word / arc queries x scene-graph factors).
 * Replaces  gather_logit_simple   src/model/joint.py:406-419:
 *   out[b][a][q][v] = sum_d txt_feat[b][q][d] * vis_feat[a][v][d];  neg_fill where !vis_mask[a][v] or !txt_mask[b][q]
 *   vis_feat [A][V][D] f32, vis_mask [A][V] bool (1 byte), txt_feat [B][Q][D] f32, txt_mask [B][Q] bool,
 *   out [B][A][Q][out_row_stride] f32 with out_row_stride >= V (= V for the reference's dense layout; a multiple of 8
 *   keeps every 128-byte store sector-aligned, which is ~2x faster -- V = 1369 is odd; elements >= V are not written)
 * One pass: bf16 tcgen05 MMAs (split = 3: hi/lo operand split, fp32-class result; split = 1: plain bf16), masks fused
 * into the store.  neg_fill is the reference's -INF = -1e20 (src/__init__.py:110, bound at joint.py:16).
 * workspace: vlgae_align_workspace_bytes(A, V, B, Q, D) bytes (packed bf16 operand tiles).
 */
size_t vlgae_align_workspace_bytes(int A, int V, int B, int Q, int D);
int vlgae_align_logits(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                       const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill, int split,
                       float *out, int out_row_stride, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Backward of vlgae_align_logits: the reference's attmap is an autograd node (einsum + masked_fill_, joint.py:413-418).
 *   grad_out [B][A][Q][grad_row_stride] (>= V) = d loss / d out;  masked entries pass no gradient:
 *   grad_txt[b][q][:] = txt_mask[b][q] * sum_{a,v} grad_out[b][a][q][v] * vis_mask[a][v] * vis_feat[a][v][:]
 *   grad_vis[a][v][:] = vis_mask[a][v] * sum_{b,q} grad_out[b][a][q][v] * txt_mask[b][q] * txt_feat[b][q][:]
 * Either output may be NULL.  tcgen05 kernels streaming grad_out once each (converted to split bf16 on the fly).
 * workspace: vlgae_align_workspace_bytes(A, V, B, Q, D) bytes.
 */
int vlgae_align_logits_backward(const float *grad_out, int grad_row_stride, const float *vis_feat,
                                const unsigned char *vis_mask, const float *txt_feat, const unsigned char *txt_mask, int A,
                                int V, int B, int Q, int D, int split, float *grad_vis, float *grad_txt, void *workspace,
                                size_t workspace_bytes, void *stream);

/*
 * Maximum of the alignment scores over the factors, without materialising the [B][A][Q][V] tensor.
 * Replaces the first half of  gather_logit_reduced   src/model/joint.py:421-432
 *   (attmap = gather_logit_simple(...); maxatt = attmap.max(dim=-1).values):
 *   maxv[b][a][q] = max_v out[b][a][q][v]   with out as defined for vlgae_align_logits (masks included), bit-identical
 *                   to the maximum of the materialised tensor;
 *   argv[b][a][q] = the smallest v that attains it (int32; may be NULL) -- the index the backward of max routes to.
 * The same tcgen05 pipeline; the epilogue reduces each 128-factor tile per query row and merges tiles with atomicMax.
 * workspace: vlgae_align_reduce_workspace_bytes(A, V, B, Q, D) bytes.
 */
size_t vlgae_align_reduce_workspace_bytes(int A, int V, int B, int Q, int D);
int vlgae_align_max_over_factors(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                                 const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill,
                                 int split, float *maxv, int *argv, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Fused grounding consumers: everything the reference's loss and decode read from attmap, without
 * the [B][A][Q][V] tensor.
 *
 * vlgae_align_maxima: BOTH maxima in one pass of the tcgen05 kernel.
 *   maxv[b][a][q] / argv  = attmap.max("V")   (joint.py:473 loss txt2vis, :520 decode)        as vlgae_align_max_over_factors
 *   maxq[b][a][v] / argq  = attmap.max("Q")   (joint.py:480 loss vis2txt)                     needs Q <= 128; either may be NULL
 *   masks as in vlgae_align_logits (a masked entry is neg_fill); arg = smallest attaining index (torch.max).
 *   workspace: vlgae_align_reduce_workspace_bytes.
 * vlgae_align_diagonal: the diagonal slab attmap[b, b] -> out [B][Q][V] (A = B; joint.py:466-469, 522-524), exact fp32.
 * vlgae_grounding_ce (A = B): out2[0] = -sum_{b,q} log_softmax_A(maxv)[b,b,q] * txt_marginal[b,q]   (joint.py:476-477)
 *                             out2[1] = -sum_{a,v} log_softmax_B(maxq)[a,a,v] * vis_mask[a,v]       (joint.py:481-483)
 *   (maxq NULL: only out2[0]).  The caller patches the POS prior into the diagonals of maxv / maxq first (joint.py:446-470).
 * vlgae_topk_rows: idx[row][0..k) = indices of the k <= 8 largest entries of x[row][0..V), descending, smaller index first
 *   on ties (match_logit.argsort(-1, descending=True)[..., :5], joint.py:594).
 * vlgae_align_max_over_factors_backward: backward of maxv through the contraction (torch routes the gradient of
 *   attmap.max to the arg-max entry): grad_txt[b,q,:] = sum_a g[b,a,q] vis[a,argv,:], grad_vis[a,argv,:] += g[b,a,q] txt[b,q,:],
 *   nothing where the arg-max entry is masked.  Either output may be NULL.  D <= 128.
 */
int vlgae_align_maxima(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                       const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill, int split,
                       float *maxv, int *argv, float *maxq, int *argq, void *workspace, size_t workspace_bytes, void *stream);
int vlgae_align_diagonal(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                         const unsigned char *txt_mask, int B, int V, int Q, int D, float neg_fill, float *out, void *stream);
int vlgae_grounding_ce(const float *maxv, const float *maxq, const float *txt_marginal, const unsigned char *vis_mask, int B,
                       int Q, int V, float *out2, void *stream);
int vlgae_topk_rows(const float *x, long long rows, int V, int k, int *idx, void *stream);
int vlgae_align_max_over_factors_backward(const float *grad_maxv, const int *argv, const float *vis_feat,
                                          const unsigned char *vis_mask, const float *txt_feat, const unsigned char *txt_mask,
                                          int A, int V, int B, int Q, int D, float *grad_vis, float *grad_txt, void *stream);

/*
 * Construction of the DMV score tensors: the step right before the chart,
 * DiscriminativeNDMV._forward (reference src/model/ldndmv.py:184-209) on the projected operands of the rank-r scorer
 * DMVFactorizedBilinear (src/model/nn/dmv_spec.py:68-76), fused with DMV1o.merge (torch_struct/distributions.py:253-265):
 *     rule[b,h,t,d,v]  = <x1[b,h,d,v,:], x2[t,d,v,:]>, log_softmax over the vocabulary t                 (:185)
 *     attach[b,h,c,v]  = rule[b,h,token[b,c],dir(h,c),v]  (dir: LEFT if c < h, RIGHT if c > h; 0 on the diagonal;
 *                        neg_fill for every child of a head with head_mask[b,h] != 0)                    (:188-198)
 *     dec[b,h,d,v,k]   = log_softmax_k(dec_score[b,h,k,d,v])                                             (:202)
 *     root[b,c]        = log_softmax(root_score)[token[b,c]]                                             (:206-207)
 *     merged_dec [B][n+1][2][2][2], merged_attach [B][n+1][n+1][2] = merge(dec, attach, root, one, zero) (:209)
 * The [B][n][T][2][2] rule tensor is never written: one streaming log-sum-exp pass over the vocabulary, then the n
 * gathered columns are recomputed while the merged tensors are written in the layout vlgae_dmv_* load.
 *   x1 [B][n][2][2][r] = attach_scorer.project1(h_parent), x2 [T][2][2][r] = attach_scorer.project2(h_child), fp32
 *   token [B][n] int64 in [0, T);  head_mask [B][n] bytes or NULL (cfg.function_mask);  r in {4, 8, 16, 32}
 *   dec_score [B][n][2][2][2] = dec_scorer(h_parent, h_dec) (decision-major, before the permute);  root_score [T]
 *   lse [B][n][2][2], root_lse [1]: written here, read by the backward;  workspace: vlgae_dmv_scores_workspace_bytes
 * vlgae_dmv_scores_backward: gradients w.r.t. x1, x2, dec_score, root_score given the gradients w.r.t. the merged tensors
 *   (e.g. the chart's marginals); the softmax over the vocabulary is recomputed in two streaming passes, not stored.
 */
size_t vlgae_dmv_scores_workspace_bytes(int B, int n);
int vlgae_dmv_scores(const float *x1, const float *x2, const int64_t *token, const unsigned char *head_mask,
                     const float *dec_score, const float *root_score, int B, int n, int T, int r, float one, float zero,
                     float neg_fill, float *merged_dec, float *merged_attach, float *lse, float *root_lse, void *workspace,
                     size_t workspace_bytes, void *stream);
int vlgae_dmv_scores_backward(const float *x1, const float *x2, const int64_t *token, const unsigned char *head_mask,
                              const float *dec_score, const float *root_score, const float *lse, const float *root_lse,
                              const float *grad_merged_dec, const float *grad_merged_attach, int B, int n, int T, int r,
                              float *grad_x1, float *grad_x2, float *grad_dec_score, float *grad_root_score, void *workspace,
                              size_t workspace_bytes, void *stream);

/*
 * Word -> factor attention of DependencyBoxRel._forward (reference joint.py:668-673):
 *     attmap = einsum("bvd,bqd->bqv", vis_feat, txt_feat).softmax(2);  out = einsum("bqv,bvh->bqh", attmap, vis_mid)
 * one caption per CTA, the factors streamed in tiles with an online softmax (the [B][n][V] attention map is never written).
 *   vis_feat [B][V][D], txt_feat [B][n][D] (the caller drops ROOT with [:, 1:] first, joint.py:669), vis_mid [B][V][H], fp32
 *   out [B][n][H];  lse [B][n] or NULL: row log-sum-exp of the scores, what the backward needs besides out
 *   n <= 64, H <= 256, (n + 64) * D + 32 * H floats of shared memory must fit (D = 128, H = 256 do).
 * vlgae_word_attention_backward: gradients of out w.r.t. the three inputs given grad_out [B][n][H]; the probabilities are
 *   recomputed from lse.  grad_vis [B][V][D], grad_txt [B][n][D], grad_mid [B][V][H]; any may be NULL.
 */
int vlgae_word_attention(const float *vis_feat, const float *txt_feat, const float *vis_mid, int B, int V, int n, int D, int H,
                         float *out, float *lse, void *stream);
int vlgae_word_attention_backward(const float *vis_feat, const float *txt_feat, const float *vis_mid, const float *out,
                                  const float *lse, const float *grad_out, int B, int V, int n, int D, int H,
                                  float *grad_vis, float *grad_txt, float *grad_mid, void *stream);

/*
 * Visual factor features with the relation MLP collapsed.  Reference:
 * VisBoxRelSimpleEncoder.forward (src/model/vis_encoder/box_rel.py:42-52) runs rel_fc = LeakyReLU(Linear(.)) on the n^2
 * pairwise means (x_i + x_j) / 2 of the box inputs, and vis_feat_unprune (src/model/joint.py:140-179) concatenates
 * [box | rel | attr | img] along the factor axis and builds the factor mask.  The Linear is affine, so
 * W ((x_i + x_j) / 2) + b = (u_i + u_j) / 2 with u = rel_fc.linear(inputs) computed once per box by the caller.
 *   u_box, u_rel [B][n][H]: pre-activations box_fc.linear(inputs), rel_fc.linear(inputs);  u_attr [B][n][H] or NULL (use_attr)
 *   box_mask [B][n] bytes;  has_img: append encoded["box"].mean(1) as the last factor (cfg.add_image);  slope: LeakyReLU
 *   mid  [B][V][H], V = n + n^2 (+ n) (+ 1):  box: lrelu(u_box[i]);  rel (i, j) at n + i n + j: lrelu((u_rel[i] + u_rel[j]) / 2);
 *        attr: lrelu(u_attr[i]);  img: mean_i lrelu(u_box[i])
 *   mask [B][V] bytes: box_mask | box_mask[i] & box_mask[j] & j > i | box_mask | 1        (joint.py:147-172)
 * vlgae_vis_factors_backward: grad_mid [B][V][H] -> gradients of the per-box pre-activations (grad_attr NULL iff u_attr NULL).
 */
int vlgae_vis_factors(const float *u_box, const float *u_rel, const float *u_attr, const unsigned char *box_mask, int B, int n,
                      int H, int has_img, float slope, float *mid, unsigned char *mask, void *stream);
int vlgae_vis_factors_backward(const float *u_box, const float *u_rel, const float *u_attr, const float *grad_mid, int B, int n,
                               int H, int has_img, float slope, float *grad_box, float *grad_rel, float *grad_attr, void *stream);

/* out[b][...] = g[b] * in[b][...]  (inner = elements per sentence): backward of partition / max. */
int vlgae_scale_rows(const float *in, const float *g, int B, size_t inner, float *out, void *stream);

/*
 * Microbenchmarks used by bench.py to measure the roofline denominators on the box:
 * MUFU (ex2.approx.f32) ops/s and FP32-pipe (FADD) ops/s over the whole chip.
 * Each writes elapsed milliseconds and the number of operations issued.
 */
int vlgae_microbench_mufu(int iters, float *ms_host, double *ops_host, void *stream);
int vlgae_microbench_fp32(int iters, float *ms_host, double *ops_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VLGAE_B200_H_ */
