// align_kernels.cuh -- internal interface between the C ABI (c_api.cu) and the alignment kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vlgae {

struct AlignArgs {
    const uint8_t *vis_packed;     // [A][VT][2*KB chunks of 128 rows x 128 B]   bf16 hi / lo, swizzled tile images
    const uint8_t *txt_packed;     // [B][QT][2*KB chunks of 128 rows x 128 B]
    const uint32_t *txt_maskbits;  // [B][QT][4]
    const uint8_t *vis_mask;       // [A][V] bool
    float *out;                    // [B][A][Q][ldv], ldv >= V
    int ldv;
    int A, V, B, Q;
    int KB, VT, QT, nq;  // k-blocks of 64, v-tiles, q-tiles, padded queries per tile (multiple of 16, <= 128)
    int BCH, stages, out_bufs, out_rows, teams, split, debug, bulk;
    uint32_t slot_bytes;
    int tile_stride;  // floats per row of a staged output tile: 128, or 132 when the rows of `out` are not 16-byte aligned
    float neg;
    float *maxv;  // MODE 2: [B][A][Q] max over the factors
    int *argv;    // MODE 2: [B][A][Q] first arg-max, or null
    float *maxq;  // MODE 2, optional: [B][A][V] max over the queries (needs Q <= 128: one query tile)
    int *argq;    // MODE 2, optional: [B][A][V] first arg-max over the queries
    uint32_t run_bytes;  // MODE 2: shared memory of the running maxima (+ 2 KB of exchange slots when maxq is set)
    long long *prof;  // debug: per CTA 8 counters of the MMA warp (clocks waiting for captions / accumulators / issuing)
};

struct AlignPlan {
    int KB, VT, QT, nq;
    size_t tile_bytes, vis_packed_bytes, txt_packed_bytes, maskbits_bytes;
};

AlignPlan align_plan(int A, int V, int B, int Q, int D, int vstep = 128);  // vstep: factors between consecutive image tiles
size_t align_workspace_bytes(int A, int V, int B, int Q, int D);
// split = 3: bf16 hi/lo split, three MMAs per product (fp32-class); split = 1: single bf16 MMA
cudaError_t launch_align(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                         int V, int B, int Q, int D, float neg, int split, float *out, int ldv, void *workspace,
                         cudaStream_t st);
// max over the factors without materialising the logits: maxv [B][A][Q] fp32, argv [B][A][Q] int32 (first arg-max;
// masked queries: neg / 0).  Workspace: align_workspace_bytes + align_reduce_bytes.
size_t align_reduce_bytes(int A, int B, int Q);
cudaError_t align_pack_operands(const float *vis, const float *txt, const uint8_t *txt_mask, int A, int V, int B, int Q, int D,
                                void *workspace, cudaStream_t st, int vstep = 128);
// backward of the logits (align_bwd_kernels.cu): g [B][A][Q][ldg]; grad_vis [A][V][D] and / or grad_txt [B][Q][D] (null = skip);
// workspace: align_workspace_bytes
cudaError_t launch_align_backward(const float *g, int ldg, const float *vis, const uint8_t *vis_mask, const float *txt,
                                  const uint8_t *txt_mask, int A, int V, int B, int Q, int D, int split, float *grad_vis,
                                  float *grad_txt, void *workspace, cudaStream_t st);
cudaError_t launch_align_reduce(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                                int V, int B, int Q, int D, float neg, int split, float *maxv, int *argv, void *workspace,
                                cudaStream_t st);
// both maxima in one pass: + maxq [B][A][V] / argq (max over the queries; Q <= 128)
cudaError_t launch_align_maxima(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                                int V, int B, int Q, int D, float neg, int split, float *maxv, int *argv, float *maxq,
                                int *argq, void *workspace, cudaStream_t st);
// small fp32 kernels of the fused grounding consumers (align_consumers.cu)
cudaError_t launch_align_diagonal(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int B,
                                  int V, int Q, int D, float neg, float *out, cudaStream_t st);
cudaError_t launch_grounding_ce(const float *maxv, const float *maxq, const float *txt_marginal, const uint8_t *vis_mask,
                                int B, int Q, int V, float *out2, cudaStream_t st);
cudaError_t launch_topk_rows(const float *x, long long rows, int V, int k, int *idx, cudaStream_t st);
cudaError_t launch_max_over_factors_backward(const float *g, const int *argv, const float *vis, const uint8_t *vis_mask,
                                             const float *txt, const uint8_t *txt_mask, int A, int V, int B, int Q, int D,
                                             float *grad_vis, float *grad_txt, cudaStream_t st);
// word -> factor attention (word_attention.cu; joint.py:668-673)
cudaError_t launch_word_attention(const float *vis, const float *txt, const float *mid, int B, int V, int n, int D, int H,
                                  float *out, float *lse, cudaStream_t st);
cudaError_t launch_word_attention_backward(const float *vis, const float *txt, const float *mid, const float *out,
                                           const float *lse, const float *gout, int B, int V, int n, int D, int H, float *gvis,
                                           float *gtxt, float *gmid, cudaStream_t st);
// visual factor features with the relation MLP collapsed (vis_factors.cu; vis_encoder/box_rel.py:42-52, joint.py:140-179)
cudaError_t launch_vis_factors(const float *u_box, const float *u_rel, const float *u_attr, const uint8_t *box_mask, int B, int n,
                               int H, int has_img, float slope, float *mid, uint8_t *mask, cudaStream_t st);
cudaError_t launch_vis_factors_backward(const float *u_box, const float *u_rel, const float *u_attr, const float *g_mid, int B, int n,
                                        int H, int has_img, float slope, float *g_box, float *g_rel, float *g_attr, cudaStream_t st);

}  // namespace vlgae
