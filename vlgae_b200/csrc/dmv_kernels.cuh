// dmv_kernels.cuh -- internal interface between the C ABI (c_api.cu), the launch logic (dmv_launch.cu) and the DMV kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vlgae {

struct DmvArgs {
    // inputs (device)
    const float *dec;        // [B][N][2][2][2]
    const float *attach;     // [B][N][N][2]
    const int64_t *lengths;  // [B]
    int B, N;
    float mask_zero;
    // log semiring
    const float *gZ;  // [B] or null
    float *Z;         // [B]
    float *gdec;      // [B][N][2][2][2] or null
    float *gattach;   // [B][N][N][2] or null
    // max semiring
    float *best;     // [B]
    int64_t *heads;  // [B][N] or null
    float *arcs;     // [B][N][N][2] or null
    float *vgdec;    // [B][N][2][2][2] or null
    // scheduling
    void *workspace;   // per-CTA chart storage when the chart does not fit in shared memory
    size_t ws_stride;  // bytes per CTA in `workspace`
    int npass;         // 1: only `first_pass`; 2: log and max CTAs interleaved
    int first_pass;    // 0 = log, 1 = max
    int nsm;           // SM count (work-item placement)
    int nb_lo, nb_hi;  // this launch handles sentences with nb_lo <= len + 1 <= nb_hi (length buckets)
    int smem_n;        // chart positions the shared-memory layout is sized for (>= nb_hi)
    long long *prof;   // optional [8] cycle counters written by the CTA of sentence 0 (debug)
    int prof_all;      // debug (build with -DVLGAE_TIMELINE, run with VLGAE_PROF_ALL=1): prof holds 8 + 4 B more words,
                       // %globaltimer at the start / end of every work item of the frontier kernel
    int no_offsets;    // debug: log-semiring sweeps on the raw scores (no per-word offsets)
    int lin_long_from; // ... and so do sentences of at least this many words (0 = none): from ~45 words on the reference's own
                       // fp32 result is > 1e-5 from the exact one and parity is judged by the three-way rule anyway
    int lin_max_len;   // frontier schedule: sentences of at most this many words run the register-state log-semiring sweeps in the
                       // LINEAR domain (0 = never; see launch_dmv for the default and why it is length-bound)
    float retry_above; // log-semiring sweeps: repeat once with corrected offsets when |log Z'| exceeds this (0 = default)
    // gather schedule (linear-domain log semiring): redo[b] = 1 when sentence b failed the sweep's self-check; the
    // follow-up frontier launch handles exactly the sentences with only[b] != 0 (null = all)
    int *redo;         // [B] or null
    const int *only;   // [B] or null
    // gather schedule: sentences are handed out through this counter (zeroed before the launch) instead of a fixed
    // stride, so a CTA that becomes resident late (another launch of the call still holds the SM) simply takes fewer
    int *counter;      // or null
    // frontier kernel, both passes in one launch, inputs in pinned HOST memory: the log CTA of a sentence republishes
    // what it staged (dec, arc scores) in device memory and the max CTA of the same sentence takes it from there, so
    // every input byte crosses PCIe once.  share_flag[b] == share_epoch once sentence b is published.
    float *share;            // [B][share_stride] or null
    unsigned *share_flag;    // [B]
    unsigned share_epoch;
    int share_stride;        // floats per sentence: N * 8 + 4 * ncells(N)
    // host entry point: the lengths are in HOST memory, and reading one from there is a PCIe round trip in front of a CTA's
    // first score load; the host side has them for free, so up to 256 travel in the kernel parameters instead
    int n_len_inline;               // sentences b < n_len_inline take their length from len_inline[b]
    unsigned char len_inline[256];  // clamped to [0, 255] (charts of that path have <= 72 positions)
};

// passes bitmask: 1 = log semiring, 2 = max semiring
bool dmv_fits_smem(int N, int passes);
int dmv_grid_for_workspace(int B);
void dmv_set_profile_buffer(long long *buf);
long long *dmv_profile_buffer();  // nullptr unless a debug buffer was registered
cudaError_t launch_dmv(const DmvArgs &a, int passes, cudaStream_t st);
// frontier schedule (dmv_frontier.cu): chart in shared memory only; `cap` = chart positions the launch is sized for
bool dmv_frontier_fits(int cap, int passes, int smem_optin);
size_t dmv_frontier_chart_bytes(int N, int passes);  // per-CTA workspace slice when the chart is in global memory
cudaError_t launch_dmv_frontier(DmvArgs a, int passes, int cap, int threads, bool reg_state, int sm_count, int max_grid,
                                cudaStream_t st);
// gather schedule (dmv_gather.cu): throughput regime, chart in shared memory, row-major squares of up to
// DMV_GATHER_MAX_POSITIONS positions
constexpr int DMV_GATHER_MAX_POSITIONS = 41;
constexpr int DMV_GATHER_COUNTERS = 16;  // ints reserved behind the redo flags
bool dmv_gather_fits(int cap, int passes, int smem_optin);
cudaError_t launch_dmv_gather(DmvArgs a, int passes, int cap, int threads, int sm_count, cudaStream_t st);
void dmv_set_schedule(int which);  // 0 = automatic, 1 = frontier, 2 = gather
void dmv_set_linear_max_len(int words);  // frontier schedule: linear-domain sweeps up to this many words; < 0 = the default
size_t dmv_ws_slice_bytes(int N, int passes);  // what vlgae_dmv_workspace_bytes reserves per CTA (either schedule)
cudaError_t launch_merge(const float *dec, const float *attach, const float *root, int B, int n, float one, float zero,
                         float *dec_w, float *attach_w, cudaStream_t st);
cudaError_t launch_scale_rows(const float *in, const float *g, int B, size_t inner, float *out, cudaStream_t st);
// score-tensor construction (dmv_scores.cu; ldndmv.py:184-209)
size_t dmv_scores_workspace_bytes(int B, int n);
cudaError_t launch_dmv_scores(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                              const float *dec_score, const float *root_score, int B, int n, int T, int r, float one, float zero,
                              float neg_fill, float *mdec, float *mattach, float *lse, float *root_lse, void *ws, int sm_count,
                              cudaStream_t st);
cudaError_t launch_dmv_scores_backward(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                                       const float *dec_score, const float *root_score, const float *lse, const float *root_lse,
                                       const float *g_mdec, const float *g_mattach, int B, int n, int T, int r, float *g_x1,
                                       float *g_x2, float *g_dec_score, float *g_root_score, void *ws, int sm_count, cudaStream_t st);
int dmv_sm_count();
cudaError_t launch_microbench(int which, int iters, float *sink, int *grid_out, int *block_out, cudaStream_t st);

}  // namespace vlgae
