// c_api.cu -- the extern "C" boundary declared in include/vlgae_b200.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vlgae_b200.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "align_kernels.cuh"
#include "deptree_kernels.cuh"
#include "dmv_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return VLGAE_E_CUDA;
}

int check_dmv(const float *dec, const float *attach, const int64_t *lengths, int B, int N) {
    if (!dec || !attach || !lengths) return fail(VLGAE_E_INVALID, "%s", "dec, attach and lengths must be non-null");
    if (B < 0) return fail(VLGAE_E_INVALID, "%s", "B must be >= 0");
    if (N < 1 || N > VLGAE_DMV_MAX_N) return fail(VLGAE_E_INVALID, "%s", "N must be in [1, 256]");
    return VLGAE_OK;
}

// workspace = per-sentence redo flags of the gather schedule (always) | chart slices (only when N is beyond shared memory)
// (+ the work counters of the gather launches behind the flags)
size_t redo_bytes(int B) { return ((size_t)(B + vlgae::DMV_GATHER_COUNTERS) * 4 + 255) & ~(size_t)255; }

int run_dmv(vlgae::DmvArgs &a, int passes, void *workspace, size_t workspace_bytes, void *stream) {
    if (a.B == 0) return VLGAE_OK;
    const size_t rb = redo_bytes(a.B);
    if (!vlgae::dmv_fits_smem(a.N, passes)) {
        const size_t need = vlgae_dmv_workspace_bytes(a.B, a.N);
        if (!workspace || workspace_bytes < need) return fail(VLGAE_E_WORKSPACE, "%s", "workspace too small for this N");
        a.workspace = (unsigned char *)workspace + rb;
    }
    // without the flags (no workspace passed) the launch logic keeps to the frontier schedule
    a.redo = (workspace && workspace_bytes >= rb && !a.share) ? (int *)workspace : nullptr;
    a.counter = a.redo ? a.redo + a.B : nullptr;
    a.npass = passes == 3 ? 2 : 1;
    a.first_pass = passes == 2 ? 1 : 0;
    cudaError_t e = vlgae::launch_dmv(a, passes, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "dmv launch");
    return VLGAE_OK;
}

}  // namespace

extern "C" {

int vlgae_version(void) { return 1; }
int vlgae_dmv_set_profile_buffer(void *device_buf) {
    vlgae::dmv_set_profile_buffer((long long *)device_buf);
    return VLGAE_OK;
}
int vlgae_dmv_set_schedule(int which) {
    if (which < 0 || which > 2) return fail(VLGAE_E_INVALID, "%s", "schedule must be 0 (auto), 1 (frontier) or 2 (gather)");
    vlgae::dmv_set_schedule(which);
    return VLGAE_OK;
}
int vlgae_dmv_set_linear_max_len(int words) {
    vlgae::dmv_set_linear_max_len(words);
    return VLGAE_OK;
}
const char *vlgae_last_error(void) { return g_err; }

size_t vlgae_dmv_workspace_bytes(int B, int N) {
    if (B <= 0 || N < 1 || N > VLGAE_DMV_MAX_N) return 0;
    if (vlgae::dmv_fits_smem(N, 3)) return redo_bytes(B);
    return redo_bytes(B) + (size_t)vlgae::dmv_grid_for_workspace(B) * vlgae::dmv_ws_slice_bytes(N, 3);
}

int vlgae_dmv_inside_outside(const float *dec, const float *attach, const int64_t *lengths, int B, int N,
                             float mask_zero, const float *gZ, float *Z, float *gdec, float *gattach, void *workspace,
                             size_t workspace_bytes, void *stream) {
    int rc = check_dmv(dec, attach, lengths, B, N);
    if (rc) return rc;
    if (!Z) return fail(VLGAE_E_INVALID, "%s", "Z must be non-null");
    vlgae::DmvArgs a;
    memset(&a, 0, sizeof(a));
    a.dec = dec; a.attach = attach; a.lengths = lengths; a.B = B; a.N = N; a.mask_zero = mask_zero;
    a.gZ = gZ; a.Z = Z; a.gdec = gdec; a.gattach = gattach;
    return run_dmv(a, 1, workspace, workspace_bytes, stream);
}

int vlgae_dmv_viterbi(const float *dec, const float *attach, const int64_t *lengths, int B, int N, float mask_zero,
                      float *best, int64_t *heads, float *arcs, float *gdec, void *workspace, size_t workspace_bytes,
                      void *stream) {
    int rc = check_dmv(dec, attach, lengths, B, N);
    if (rc) return rc;
    if (!best) return fail(VLGAE_E_INVALID, "%s", "best must be non-null");
    vlgae::DmvArgs a;
    memset(&a, 0, sizeof(a));
    a.dec = dec; a.attach = attach; a.lengths = lengths; a.B = B; a.N = N; a.mask_zero = mask_zero;
    a.best = best; a.heads = heads; a.arcs = arcs; a.vgdec = gdec;
    return run_dmv(a, 2, workspace, workspace_bytes, stream);
}

int vlgae_dmv_parse(const float *dec, const float *attach, const int64_t *lengths, int B, int N, float mask_zero,
                    const float *gZ, float *Z, float *gdec, float *gattach, float *best, int64_t *heads, float *arcs,
                    float *vgdec, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_dmv(dec, attach, lengths, B, N);
    if (rc) return rc;
    if (!Z || !best) return fail(VLGAE_E_INVALID, "%s", "Z and best must be non-null");
    vlgae::DmvArgs a;
    memset(&a, 0, sizeof(a));
    a.dec = dec; a.attach = attach; a.lengths = lengths; a.B = B; a.N = N; a.mask_zero = mask_zero;
    a.gZ = gZ; a.Z = Z; a.gdec = gdec; a.gattach = gattach;
    a.best = best; a.heads = heads; a.arcs = arcs; a.vgdec = vgdec;
    return run_dmv(a, 3, workspace, workspace_bytes, stream);
}

// async_only: vlgae_dmv_parse_host_async -- zero-copy or nothing, and no synchronisation
static int parse_host_impl(const float *dec_host, const float *attach_host, const int64_t *lengths_host, int B, int N,
                           float mask_zero, float *Z_host, float *gdec_host, float *gattach_host, float *best_host,
                           int64_t *heads_host, void *stream, bool async_only) {
    int rc = check_dmv(dec_host, attach_host, lengths_host, B, N);
    if (rc) return rc;
    if (B == 0) return VLGAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const auto t_enter = std::chrono::steady_clock::now();
    const size_t nd = (size_t)B * N * 8, na = (size_t)B * N * N * 2;
    const size_t ws = vlgae_dmv_workspace_bytes(B, N);
    // Zero-copy path: when every host buffer is pinned (cudaHostAlloc / cudaHostRegister -- what torch's pin_memory()
    // gives), the kernel reads the potentials from and writes the results to host memory itself.  Each CTA pulls its
    // own sentence over PCIe when it starts and pushes its marginals when it ends, so the transfers of the short
    // sentences hide behind the chart sweeps of the long ones, and the eight per-call copies (each a few us of fixed
    // cost) disappear: cfg2 182 us -> see DESIGN.md.  Pageable buffers take the staged path below.
    static const bool env_zero_copy = [] { const char *v = getenv("VLGAE_ZERO_COPY"); return !(v && v[0] == '0'); }();
    if (env_zero_copy && vlgae::dmv_fits_smem(N, 3)) {
        const void *hp[8] = {dec_host, attach_host, lengths_host, Z_host, gdec_host, gattach_host, best_host, heads_host};
        void *dp[8];
        bool pinned = true;
        for (int k = 0; k < 8 && pinned; ++k) {
            dp[k] = nullptr;
            if (!hp[k]) continue;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hp[k]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                pinned = false;
            } else {
                dp[k] = at.devicePointer;
            }
        }
        // best and Z are required by the kernels: fall through to the staged path if the caller did not ask for them
        if (pinned && dp[3] && dp[6]) {
            const bool want_grad = gdec_host || gattach_host;
            // device scratch through which the log CTA of a sentence hands the staged inputs to the max CTA
            // (DmvArgs::share): every input byte crosses PCIe once instead of twice
            // one buffer per (device, stream): calls in flight on different streams (vlgae_dmv_parse_host_async) must not
            // share it; calls on one stream are serialised by the stream
            struct Share { unsigned char *buf; size_t bytes; unsigned epoch; int dev; cudaStream_t st; };
            static thread_local Share shares[4] = {};
            static thread_local int share_next = 0;
            int cur_dev = 0;
            cudaGetDevice(&cur_dev);
            Share *sh = nullptr;
            for (Share &c : shares)
                if (c.buf && c.dev == cur_dev && c.st == st) sh = &c;
            if (!sh) {
                for (Share &c : shares)
                    if (!c.buf) { sh = &c; break; }
                if (!sh) {  // every slot taken: recycle one (its owner's work is drained first)
                    sh = &shares[share_next];
                    share_next = (share_next + 1) % 4;
                    if (sh->dev == cur_dev) { cudaStreamSynchronize(sh->st); cudaFree(sh->buf); }
                    *sh = Share{};
                }
                sh->dev = cur_dev; sh->st = st;
            }
            unsigned char *&share = sh->buf;
            size_t &share_bytes = sh->bytes;
            unsigned &epoch = sh->epoch;
            const int stride = N * 8 + 4 * (N * (N + 1) / 2);
            const size_t need = (size_t)B * stride * 4 + (size_t)B * 4 + 64;
            if (share_bytes < need) {
                if (share) { cudaStreamSynchronize(st); cudaFree(share); share = nullptr; share_bytes = 0; }
                cudaError_t ea = cudaMalloc((void **)&share, need);
                if (ea != cudaSuccess) return cuda_fail(ea, "cudaMalloc");
                ea = cudaMemsetAsync(share, 0, need, st);
                if (ea != cudaSuccess) return cuda_fail(ea, "cudaMemset");
                share_bytes = need;
                epoch = 0;
            }
            vlgae::DmvArgs a;
            memset(&a, 0, sizeof(a));
            a.dec = (const float *)dp[0]; a.attach = (const float *)dp[1]; a.lengths = (const int64_t *)dp[2];
            a.B = B; a.N = N; a.mask_zero = mask_zero;
            a.Z = (float *)dp[3]; a.gdec = want_grad ? (float *)dp[4] : nullptr; a.gattach = want_grad ? (float *)dp[5] : nullptr;
            a.best = (float *)dp[6]; a.heads = (int64_t *)dp[7];
            a.share = (float *)share;
            a.share_flag = (unsigned *)(share + (((size_t)B * stride * 4 + 15) & ~(size_t)15));
            a.share_epoch = ++epoch;
            a.share_stride = stride;
            if (B <= (int)sizeof(a.len_inline)) {  // the lengths are host-readable: hand them over by value
                for (int b = 0; b < B; ++b) {
                    const int64_t l = lengths_host[b];
                    a.len_inline[b] = (unsigned char)(l < 0 ? 0 : (l > 255 ? 255 : l));
                }
                a.n_len_inline = B;
            }
            static const bool trace = getenv("VLGAE_E2E_TRACE") != nullptr;  // host-side breakdown of the call (debug)
            const auto t_launch = std::chrono::steady_clock::now();
            rc = run_dmv(a, 3, nullptr, 0, stream);
            if (rc) return rc;
            if (async_only) return VLGAE_OK;  // results are in host memory once `stream` has drained
            const auto t_issued = std::chrono::steady_clock::now();
            cudaError_t es = cudaStreamSynchronize(st);
            if (trace) {
                const auto t_end = std::chrono::steady_clock::now();
                auto us = [](auto d) { return std::chrono::duration<double, std::micro>(d).count(); };
                fprintf(stderr, "[vlgae] parse_host: prepare %.1f us, launch %.1f us, wait %.1f us\n", us(t_launch - t_enter),
                        us(t_issued - t_launch), us(t_end - t_issued));
            }
            return es == cudaSuccess ? VLGAE_OK : cuda_fail(es, "sync");
        }
    }
    if (async_only)
        return fail(VLGAE_E_INVALID, "%s", "vlgae_dmv_parse_host_async needs pinned host buffers (Z and best included) and N within shared memory");
    // one device arena: dec | attach | gdec | gattach | Z | best | lengths | heads | workspace
    const size_t fl = nd + na + nd + na + 2 * (size_t)B;
    const size_t bytes = ((fl * 4 + 15) & ~(size_t)15) + (size_t)B * 8 + (size_t)B * N * 8 + 256 + ws;
    // grow-only device arena kept for the life of the process (one per host thread): the end-to-end call must not
    // pay an allocation per batch
    static thread_local unsigned char *arena = nullptr;
    static thread_local size_t arena_bytes = 0;
    static thread_local int arena_dev = -1;
    cudaError_t e = cudaSuccess;
    {
        int cur_dev = 0;
        cudaGetDevice(&cur_dev);
        if (arena && arena_dev != cur_dev) { arena = nullptr; arena_bytes = 0; }
        arena_dev = cur_dev;
    }
    if (arena_bytes < bytes) {
        if (arena) { cudaStreamSynchronize(st); cudaFree(arena); arena = nullptr; arena_bytes = 0; }
        e = cudaMalloc((void **)&arena, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        arena_bytes = bytes;
    }
    float *d_dec = (float *)arena, *d_att = d_dec + nd, *d_gdec = d_att + na, *d_gatt = d_gdec + nd;
    float *d_Z = d_gatt + na, *d_best = d_Z + B;
    int64_t *d_len = (int64_t *)(arena + ((fl * 4 + 15) & ~(size_t)15));
    int64_t *d_heads = d_len + B;
    void *d_ws = (void *)(((uintptr_t)(d_heads + (size_t)B * N) + 255) & ~(uintptr_t)255);
#define CK(x, w) do { e = (x); if (e != cudaSuccess) return cuda_fail(e, w); } while (0)
    CK(cudaMemcpyAsync(d_dec, dec_host, nd * 4, cudaMemcpyHostToDevice, st), "H2D dec");
    CK(cudaMemcpyAsync(d_att, attach_host, na * 4, cudaMemcpyHostToDevice, st), "H2D attach");
    CK(cudaMemcpyAsync(d_len, lengths_host, (size_t)B * 8, cudaMemcpyHostToDevice, st), "H2D lengths");
    const bool want_grad = gdec_host || gattach_host;
    rc = vlgae_dmv_parse(d_dec, d_att, d_len, B, N, mask_zero, nullptr, d_Z, want_grad ? d_gdec : nullptr,
                         want_grad ? d_gatt : nullptr, d_best, d_heads, nullptr, nullptr, d_ws, ws, stream);
    if (rc) return rc;
    if (Z_host) CK(cudaMemcpyAsync(Z_host, d_Z, (size_t)B * 4, cudaMemcpyDeviceToHost, st), "D2H Z");
    if (best_host) CK(cudaMemcpyAsync(best_host, d_best, (size_t)B * 4, cudaMemcpyDeviceToHost, st), "D2H best");
    if (gdec_host) CK(cudaMemcpyAsync(gdec_host, d_gdec, nd * 4, cudaMemcpyDeviceToHost, st), "D2H gdec");
    if (gattach_host) CK(cudaMemcpyAsync(gattach_host, d_gatt, na * 4, cudaMemcpyDeviceToHost, st), "D2H gattach");
    if (heads_host) CK(cudaMemcpyAsync(heads_host, d_heads, (size_t)B * N * 8, cudaMemcpyDeviceToHost, st), "D2H heads");
    CK(cudaStreamSynchronize(st), "sync");
#undef CK
    return VLGAE_OK;
}

int vlgae_dmv_parse_host(const float *dec_host, const float *attach_host, const int64_t *lengths_host, int B, int N,
                         float mask_zero, float *Z_host, float *gdec_host, float *gattach_host, float *best_host,
                         int64_t *heads_host, void *stream) {
    return parse_host_impl(dec_host, attach_host, lengths_host, B, N, mask_zero, Z_host, gdec_host, gattach_host, best_host,
                           heads_host, stream, false);
}

int vlgae_dmv_parse_host_async(const float *dec_host, const float *attach_host, const int64_t *lengths_host, int B, int N,
                               float mask_zero, float *Z_host, float *gdec_host, float *gattach_host, float *best_host,
                               int64_t *heads_host, void *stream) {
    if (!Z_host || !best_host) return fail(VLGAE_E_INVALID, "%s", "Z_host and best_host must be non-null");
    return parse_host_impl(dec_host, attach_host, lengths_host, B, N, mask_zero, Z_host, gdec_host, gattach_host, best_host,
                           heads_host, stream, true);
}

int vlgae_dmv_merge(const float *dec, const float *attach, const float *root, int B, int n, float one, float zero,
                    float *dec_w, float *attach_w, void *stream) {
    if (!dec || !attach || !root || !dec_w || !attach_w) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B < 0 || n < 0 || n + 1 > VLGAE_DMV_MAX_N) return fail(VLGAE_E_INVALID, "%s", "bad B or n");
    if (B == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_merge(dec, attach, root, B, n, one, zero, dec_w, attach_w, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "merge launch");
}

int vlgae_scale_rows(const float *in, const float *g, int B, size_t inner, float *out, void *stream) {
    if (!in || !g || !out) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B <= 0 || inner == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_scale_rows(in, g, B, inner, out, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "scale_rows launch");
}

// DependencyCRF runs on the DMV kernels (see deptree_kernels.cu): workspace = dec | attach2 | g2 | DMV workspace.
static size_t deptree_dmv_bytes(int B, int N, size_t *o_att, size_t *o_g2, size_t *o_ws) {
    const size_t dec = (((size_t)B * N * 8 * 4) + 255) & ~(size_t)255, att = (((size_t)B * N * N * 2 * 4) + 255) & ~(size_t)255;
    if (o_att) *o_att = dec;
    if (o_g2) *o_g2 = dec + att;
    if (o_ws) *o_ws = dec + 2 * att;
    return dec + 2 * att + vlgae_dmv_workspace_bytes(B, N) + 256;
}

size_t vlgae_deptree_workspace_bytes(int B, int N) {
    if (B <= 0 || N < 1 || N > VLGAE_DMV_MAX_N) return 0;
    return deptree_dmv_bytes(B, N, nullptr, nullptr, nullptr);
}

int vlgae_deptree(const float *arc, const int64_t *lengths, int B, int N, float fill, float mask_zero, int semiring,
                  float *out, float *marginals, int64_t *heads, void *workspace, size_t workspace_bytes, void *stream) {
    if (!arc || !lengths || !out) return fail(VLGAE_E_INVALID, "%s", "arc, lengths and out must be non-null");
    if (B < 0 || N < 1 || N > VLGAE_DMV_MAX_N) return fail(VLGAE_E_INVALID, "%s", "bad B or N");
    if (semiring != 0 && semiring != 1) return fail(VLGAE_E_INVALID, "%s", "semiring must be 0 (log) or 1 (max)");
    if (B == 0) return VLGAE_OK;
    if (!workspace || workspace_bytes < vlgae_deptree_workspace_bytes(B, N))
        return fail(VLGAE_E_WORKSPACE, "%s", "deptree workspace too small");
    (void)fill;  // positions beyond lengths[b] never enter the chart: every sentence is swept at its own length
    size_t o_att, o_g2, o_ws;
    deptree_dmv_bytes(B, N, &o_att, &o_g2, &o_ws);
    unsigned char *w = reinterpret_cast<unsigned char *>(workspace);
    float *dec = reinterpret_cast<float *>(w), *att = reinterpret_cast<float *>(w + o_att), *g2 = reinterpret_cast<float *>(w + o_g2);
    void *dws = reinterpret_cast<void *>(((uintptr_t)(w + o_ws) + 255) & ~(uintptr_t)255);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = vlgae::launch_deptree_expand(arc, B, N, att, dec, st);
    if (e != cudaSuccess) return cuda_fail(e, "deptree expand");
    vlgae::DmvArgs a;
    memset(&a, 0, sizeof(a));
    a.dec = dec; a.attach = att; a.lengths = lengths; a.B = B; a.N = N; a.mask_zero = mask_zero;
    int rc;
    if (semiring == 0) {
        a.Z = out; a.gattach = marginals ? g2 : nullptr;
        rc = run_dmv(a, 1, dws, vlgae_dmv_workspace_bytes(B, N), stream);
    } else {
        a.best = out; a.heads = heads; a.arcs = marginals ? g2 : nullptr;
        rc = run_dmv(a, 2, dws, vlgae_dmv_workspace_bytes(B, N), stream);
    }
    if (rc) return rc;
    if (marginals) {
        e = vlgae::launch_deptree_collapse(g2, B, N, marginals, st);
        if (e != cudaSuccess) return cuda_fail(e, "deptree collapse");
    }
    return VLGAE_OK;
}

size_t vlgae_align_workspace_bytes(int A, int V, int B, int Q, int D) {
    if (A <= 0 || V <= 0 || B <= 0 || Q <= 0 || D <= 0 || D > VLGAE_ALIGN_MAX_D) return 0;
    return vlgae::align_workspace_bytes(A, V, B, Q, D);
}

int vlgae_align_logits(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                       const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill, int split,
                       float *out, int out_row_stride, void *workspace, size_t workspace_bytes, void *stream) {
    if (!vis_feat || !vis_mask || !txt_feat || !txt_mask || !out) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (A < 0 || V < 0 || B < 0 || Q < 0) return fail(VLGAE_E_INVALID, "%s", "negative extent");
    if (D < 1 || D > VLGAE_ALIGN_MAX_D) return fail(VLGAE_E_INVALID, "%s", "D must be in [1, 128]");
    if (split != 1 && split != 3) return fail(VLGAE_E_INVALID, "%s", "split must be 1 or 3");
    if (out_row_stride < V) return fail(VLGAE_E_INVALID, "%s", "out_row_stride must be >= V");
    if (A == 0 || V == 0 || B == 0 || Q == 0) return VLGAE_OK;
    const size_t need = vlgae::align_workspace_bytes(A, V, B, Q, D);
    if (!workspace || workspace_bytes < need) return fail(VLGAE_E_WORKSPACE, "%s", "alignment workspace too small");
    cudaError_t e = vlgae::launch_align(vis_feat, vis_mask, txt_feat, txt_mask, A, V, B, Q, D, neg_fill, split, out,
                                        out_row_stride, workspace, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "align launch");
}

int vlgae_align_logits_backward(const float *grad_out, int grad_row_stride, const float *vis_feat,
                                const unsigned char *vis_mask, const float *txt_feat, const unsigned char *txt_mask, int A,
                                int V, int B, int Q, int D, int split, float *grad_vis, float *grad_txt, void *workspace,
                                size_t workspace_bytes, void *stream) {
    if (!grad_out || !vis_feat || !vis_mask || !txt_feat || !txt_mask) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (A < 0 || V < 0 || B < 0 || Q < 0) return fail(VLGAE_E_INVALID, "%s", "negative extent");
    if (D < 1 || D > VLGAE_ALIGN_MAX_D) return fail(VLGAE_E_INVALID, "%s", "D must be in [1, 128]");
    if (split != 1 && split != 3) return fail(VLGAE_E_INVALID, "%s", "split must be 1 or 3");
    if (grad_row_stride < V) return fail(VLGAE_E_INVALID, "%s", "grad_row_stride must be >= V");
    if (A == 0 || V == 0 || B == 0 || Q == 0 || (!grad_vis && !grad_txt)) return VLGAE_OK;
    const size_t need = vlgae::align_workspace_bytes(A, V, B, Q, D);
    if (!workspace || workspace_bytes < need) return fail(VLGAE_E_WORKSPACE, "%s", "alignment workspace too small");
    cudaError_t e = vlgae::launch_align_backward(grad_out, grad_row_stride, vis_feat, vis_mask, txt_feat, txt_mask, A, V, B, Q,
                                                 D, split, grad_vis, grad_txt, workspace, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "align backward launch");
}

size_t vlgae_align_reduce_workspace_bytes(int A, int V, int B, int Q, int D) {
    const size_t base = vlgae_align_workspace_bytes(A, V, B, Q, D);
    if (base == 0) return 0;
    return ((base + 255) & ~(size_t)255) + 256 + vlgae::align_reduce_bytes(A, B, Q);
}

int vlgae_align_max_over_factors(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                                 const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill,
                                 int split, float *maxv, int *argv, void *workspace, size_t workspace_bytes, void *stream) {
    if (!vis_feat || !vis_mask || !txt_feat || !txt_mask || !maxv) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (A < 0 || V < 0 || B < 0 || Q < 0) return fail(VLGAE_E_INVALID, "%s", "negative extent");
    if (D < 1 || D > VLGAE_ALIGN_MAX_D) return fail(VLGAE_E_INVALID, "%s", "D must be in [1, 128]");
    if (split != 1 && split != 3) return fail(VLGAE_E_INVALID, "%s", "split must be 1 or 3");
    if (A == 0 || B == 0 || Q == 0) return VLGAE_OK;
    if (V == 0) return fail(VLGAE_E_INVALID, "%s", "max over an empty factor axis");
    const size_t need = vlgae_align_reduce_workspace_bytes(A, V, B, Q, D);
    if (!workspace || workspace_bytes < need) return fail(VLGAE_E_WORKSPACE, "%s", "alignment workspace too small");
    cudaError_t e = vlgae::launch_align_reduce(vis_feat, vis_mask, txt_feat, txt_mask, A, V, B, Q, D, neg_fill, split, maxv,
                                               argv, workspace, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "align reduce launch");
}

int vlgae_align_maxima(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                       const unsigned char *txt_mask, int A, int V, int B, int Q, int D, float neg_fill, int split,
                       float *maxv, int *argv, float *maxq, int *argq, void *workspace, size_t workspace_bytes, void *stream) {
    if (!vis_feat || !vis_mask || !txt_feat || !txt_mask || !maxv) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (A < 0 || V < 0 || B < 0 || Q < 0) return fail(VLGAE_E_INVALID, "%s", "negative extent");
    if (D < 1 || D > VLGAE_ALIGN_MAX_D) return fail(VLGAE_E_INVALID, "%s", "D must be in [1, 128]");
    if (split != 1 && split != 3) return fail(VLGAE_E_INVALID, "%s", "split must be 1 or 3");
    if (maxq && Q > 128) return fail(VLGAE_E_INVALID, "%s", "the fused max over the queries needs Q <= 128");
    if (A == 0 || B == 0 || Q == 0) return VLGAE_OK;
    if (V == 0) return fail(VLGAE_E_INVALID, "%s", "max over an empty factor axis");
    const size_t need = vlgae_align_reduce_workspace_bytes(A, V, B, Q, D);
    if (!workspace || workspace_bytes < need) return fail(VLGAE_E_WORKSPACE, "%s", "alignment workspace too small");
    cudaError_t e = vlgae::launch_align_maxima(vis_feat, vis_mask, txt_feat, txt_mask, A, V, B, Q, D, neg_fill, split, maxv, argv,
                                               maxq, argq, workspace, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "align maxima launch");
}

int vlgae_align_diagonal(const float *vis_feat, const unsigned char *vis_mask, const float *txt_feat,
                         const unsigned char *txt_mask, int B, int V, int Q, int D, float neg_fill, float *out, void *stream) {
    if (!vis_feat || !vis_mask || !txt_feat || !txt_mask || !out) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B < 0 || V < 0 || Q < 0 || D < 1 || D > 256) return fail(VLGAE_E_INVALID, "%s", "bad extent");
    if (B == 0 || V == 0 || Q == 0) return VLGAE_OK;
    if (((size_t)((D + 3) & ~3) * (64 + (size_t)Q)) * 4 > 200 * 1024) return fail(VLGAE_E_INVALID, "%s", "Q * D too large for the diagonal kernel");
    cudaError_t e = vlgae::launch_align_diagonal(vis_feat, vis_mask, txt_feat, txt_mask, B, V, Q, D, neg_fill, out, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "align diagonal launch");
}

int vlgae_grounding_ce(const float *maxv, const float *maxq, const float *txt_marginal, const unsigned char *vis_mask, int B,
                       int Q, int V, float *out2, void *stream) {
    if (!maxv || !txt_marginal || !out2 || (maxq && !vis_mask)) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B <= 0 || Q <= 0 || (maxq && V <= 0)) return fail(VLGAE_E_INVALID, "%s", "bad extent");
    cudaError_t e = vlgae::launch_grounding_ce(maxv, maxq, txt_marginal, vis_mask, B, Q, V, out2, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "grounding ce launch");
}

int vlgae_topk_rows(const float *x, long long rows, int V, int k, int *idx, void *stream) {
    if (!x || !idx) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (rows < 0 || V < 1 || k < 1 || k > 8) return fail(VLGAE_E_INVALID, "%s", "rows >= 0, V >= 1, 1 <= k <= 8 expected");
    if (rows == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_topk_rows(x, rows, V, k, idx, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "topk launch");
}

int vlgae_align_max_over_factors_backward(const float *grad_maxv, const int *argv, const float *vis_feat,
                                          const unsigned char *vis_mask, const float *txt_feat, const unsigned char *txt_mask,
                                          int A, int V, int B, int Q, int D, float *grad_vis, float *grad_txt, void *stream) {
    if (!grad_maxv || !argv || !vis_feat || !vis_mask || !txt_feat || !txt_mask) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (A < 0 || V < 0 || B < 0 || Q < 0 || D < 1 || D > VLGAE_ALIGN_MAX_D) return fail(VLGAE_E_INVALID, "%s", "bad extent");
    if (A == 0 || V == 0 || B == 0 || Q == 0 || (!grad_vis && !grad_txt)) return VLGAE_OK;
    cudaError_t e = vlgae::launch_max_over_factors_backward(grad_maxv, argv, vis_feat, vis_mask, txt_feat, txt_mask, A, V, B, Q, D,
                                                            grad_vis, grad_txt, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "max backward launch");
}

size_t vlgae_dmv_scores_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 256;
    return vlgae::dmv_scores_workspace_bytes(B, n);
}

static int check_scores(int B, int n, int T, int r, size_t workspace_bytes, const void *workspace) {
    if (B < 0 || n < 1 || n + 1 > VLGAE_DMV_MAX_N || T < 1) return fail(VLGAE_E_INVALID, "%s", "B >= 0, 1 <= n <= 255, T >= 1 expected");
    if (r != 4 && r != 8 && r != 16 && r != 32) return fail(VLGAE_E_INVALID, "%s", "rank must be 4, 8, 16 or 32");
    if (B > 0 && (!workspace || workspace_bytes < vlgae::dmv_scores_workspace_bytes(B, n)))
        return fail(VLGAE_E_WORKSPACE, "%s", "workspace too small (vlgae_dmv_scores_workspace_bytes)");
    return VLGAE_OK;
}

int vlgae_dmv_scores(const float *x1, const float *x2, const int64_t *token, const unsigned char *head_mask,
                     const float *dec_score, const float *root_score, int B, int n, int T, int r, float one, float zero,
                     float neg_fill, float *merged_dec, float *merged_attach, float *lse, float *root_lse, void *workspace,
                     size_t workspace_bytes, void *stream) {
    if (!x1 || !x2 || !token || !dec_score || !root_score || !merged_dec || !merged_attach || !lse || !root_lse)
        return fail(VLGAE_E_INVALID, "%s", "null pointer");
    int rc = check_scores(B, n, T, r, workspace_bytes, workspace);
    if (rc) return rc;
    if (B == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_dmv_scores(x1, x2, (const long long *)token, head_mask, dec_score, root_score, B, n, T, r, one,
                                             zero, neg_fill, merged_dec, merged_attach, lse, root_lse, workspace,
                                             vlgae::dmv_sm_count(), (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "dmv scores launch");
}

int vlgae_dmv_scores_backward(const float *x1, const float *x2, const int64_t *token, const unsigned char *head_mask,
                              const float *dec_score, const float *root_score, const float *lse, const float *root_lse,
                              const float *grad_merged_dec, const float *grad_merged_attach, int B, int n, int T, int r,
                              float *grad_x1, float *grad_x2, float *grad_dec_score, float *grad_root_score, void *workspace,
                              size_t workspace_bytes, void *stream) {
    if (!x1 || !x2 || !token || !dec_score || !root_score || !lse || !root_lse || !grad_merged_dec || !grad_merged_attach ||
        !grad_x1 || !grad_x2 || !grad_dec_score || !grad_root_score)
        return fail(VLGAE_E_INVALID, "%s", "null pointer");
    int rc = check_scores(B, n, T, r, workspace_bytes, workspace);
    if (rc) return rc;
    if (B == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_dmv_scores_backward(x1, x2, (const long long *)token, head_mask, dec_score, root_score, lse, root_lse,
                                                      grad_merged_dec, grad_merged_attach, B, n, T, r, grad_x1, grad_x2,
                                                      grad_dec_score, grad_root_score, workspace, vlgae::dmv_sm_count(),
                                                      (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "dmv scores backward launch");
}

int vlgae_word_attention(const float *vis_feat, const float *txt_feat, const float *vis_mid, int B, int V, int n, int D, int H,
                         float *out, float *lse, void *stream) {
    if (!vis_feat || !txt_feat || !vis_mid || !out) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B < 0 || V < 1 || n < 0 || n > 64 || D < 1 || H < 1 || H > 256)
        return fail(VLGAE_E_INVALID, "%s", "B >= 0, V >= 1, 0 <= n <= 64, D >= 1, 1 <= H <= 256 expected");
    if (B == 0 || n == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_word_attention(vis_feat, txt_feat, vis_mid, B, V, n, D, H, out, lse, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "word attention launch");
}

int vlgae_word_attention_backward(const float *vis_feat, const float *txt_feat, const float *vis_mid, const float *out,
                                  const float *lse, const float *grad_out, int B, int V, int n, int D, int H,
                                  float *grad_vis, float *grad_txt, float *grad_mid, void *stream) {
    if (!vis_feat || !txt_feat || !vis_mid || !out || !lse || !grad_out) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B < 0 || V < 1 || n < 1 || n > 64 || D < 1 || H < 1 || H > 256)
        return fail(VLGAE_E_INVALID, "%s", "B >= 0, V >= 1, 1 <= n <= 64, D >= 1, 1 <= H <= 256 expected");
    if (B == 0 || (!grad_vis && !grad_txt && !grad_mid)) return VLGAE_OK;
    cudaError_t e = vlgae::launch_word_attention_backward(vis_feat, txt_feat, vis_mid, out, lse, grad_out, B, V, n, D, H, grad_vis,
                                                          grad_txt, grad_mid, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "word attention backward launch");
}

int vlgae_vis_factors(const float *u_box, const float *u_rel, const float *u_attr, const unsigned char *box_mask, int B, int n,
                      int H, int has_img, float slope, float *mid, unsigned char *mask, void *stream) {
    if (!u_box || !u_rel || !box_mask || !mid || !mask) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if (B < 0 || n < 1 || n > 1024 || H < 1) return fail(VLGAE_E_INVALID, "%s", "B >= 0, 1 <= n <= 1024, H >= 1 expected");
    if (B == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_vis_factors(u_box, u_rel, u_attr, box_mask, B, n, H, has_img, slope, mid, mask, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "vis factors launch");
}

int vlgae_vis_factors_backward(const float *u_box, const float *u_rel, const float *u_attr, const float *grad_mid, int B, int n,
                               int H, int has_img, float slope, float *grad_box, float *grad_rel, float *grad_attr, void *stream) {
    if (!u_box || !u_rel || !grad_mid || !grad_box || !grad_rel) return fail(VLGAE_E_INVALID, "%s", "null pointer");
    if ((u_attr == nullptr) != (grad_attr == nullptr)) return fail(VLGAE_E_INVALID, "%s", "grad_attr must be given exactly when u_attr is");
    if (B < 0 || n < 1 || n > 1024 || H < 1) return fail(VLGAE_E_INVALID, "%s", "B >= 0, 1 <= n <= 1024, H >= 1 expected");
    if (B == 0) return VLGAE_OK;
    cudaError_t e = vlgae::launch_vis_factors_backward(u_box, u_rel, u_attr, grad_mid, B, n, H, has_img, slope, grad_box, grad_rel,
                                                       grad_attr, (cudaStream_t)stream);
    return e == cudaSuccess ? VLGAE_OK : cuda_fail(e, "vis factors backward launch");
}

static int microbench(int which, int iters, float *ms_host, double *ops_host, void *stream) {
    if (!ms_host || !ops_host || iters < 1) return fail(VLGAE_E_INVALID, "%s", "bad microbench arguments");
    cudaStream_t st = (cudaStream_t)stream;
    float *sink = nullptr;
    cudaError_t e = cudaMalloc((void **)&sink, 4);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = 0, block = 0;
    vlgae::launch_microbench(which, 16, sink, &grid, &block, st);  // warm-up
    cudaEventRecord(a, st);
    e = vlgae::launch_microbench(which, iters, sink, &grid, &block, st);
    cudaEventRecord(b, st);
    cudaError_t e2 = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    if (e != cudaSuccess) return cuda_fail(e, "microbench launch");
    if (e2 != cudaSuccess) return cuda_fail(e2, "microbench sync");
    *ms_host = ms;
    *ops_host = (double)grid * block * (double)iters * 64.0;
    return VLGAE_OK;
}
int vlgae_microbench_mufu(int iters, float *ms_host, double *ops_host, void *stream) { return microbench(0, iters, ms_host, ops_host, stream); }
int vlgae_microbench_fp32(int iters, float *ms_host, double *ops_host, void *stream) { return microbench(1, iters, ms_host, ops_host, stream); }

}  // extern "C"
