// dmv_scores.cu -- construction of the DMV score tensors, the step right before the chart.
//
// Reference: DiscriminativeNDMV._forward, /root/reference/src/model/ldndmv.py:184-209, with the rank-r bilinear scorer
// DMVFactorizedBilinear.forward (/root/reference/src/model/nn/dmv_spec.py:68-76) and DMV1o.merge
// (/root/reference/src/model/torch_struct/distributions.py:253-265):
//
//     attach_rule[b,h,t,d,v] = <x1[b,h,d,v,:], x2[t,d,v,:]>                 (einsum "bhdve,bcdve->bhcdv", x2 broadcast over b)
//     attach_rule = attach_rule.log_softmax(2)                              over the vocabulary t            (:185)
//     attach[b,h,c,v] = attach_rule[b,h,token[b,c],dir(h,c),v], 0 on the diagonal, -INF for masked heads     (:188-198)
//     dec[b,h,d,v,:]  = dec_score[b,h,:,d,v].log_softmax over the 2 decisions                                (:202)
//     root[b,c]       = root_score.log_softmax(-1)[token[b,c]]                                               (:206-207)
//     merged_dec, merged_attach = DMV1o.merge(dec, attach, root)                                             (:209)
//
// The reference materialises attach_rule [B,n,T,2,2] (0.8 GB at B=128, n=40, T=10k) to read n of its T columns.  Here the
// log-sum-exp over the vocabulary is one streaming pass (x2 tiles in shared memory, one thread per (b,h,d,v) row, a warp
// = 32 rows of one (d,v) so every operand read is a broadcast), and the n gathered columns are recomputed as r-term dot
// products while the merged tensors are written -- directly in the layout the chart kernels load.
// Backward (given the gradients w.r.t. the merged tensors, i.e. the chart's marginals): the softmax is recomputed in two
// more streaming passes, one row-major (d x1) and one token-major (d x2), instead of being stored.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmv_kernels.cuh"

namespace vlgae {
namespace {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int SC_TT = 64;  // tokens per shared-memory tile (row pass) / (b,h) pairs per tile (token pass)

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// online log-sum-exp state (m, s): value = m + log(s)
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
    const float mx = fmaxf(m, m2);
    if (mx == -INFINITY) return;  // both empty
    s = s * ex2f((m - mx) * LOG2E_F) + s2 * ex2f((m2 - mx) * LOG2E_F);
    m = mx;
}

// ---- pass 1: partial log-sum-exp over a slice of the vocabulary --------------------------------------------------
// grid (ceil(B n / 32), tsplit); 128 threads: warp = (d, v), lane = (b, h) pair.  part[y][pair][dv] = (m, s).
template <int R>
__global__ void __launch_bounds__(128) scores_lse_kernel(const float *__restrict__ x1, const float *__restrict__ x2, int npair,
                                                         int T, int tchunk, float2 *__restrict__ part) {
    __shared__ __align__(16) float tile[SC_TT * 4 * R];
    const int dv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 32 + lane;
    const bool live = pair < npair;
    float a[R];
#pragma unroll
    for (int e = 0; e < R; ++e) a[e] = live ? x1[((size_t)pair * 4 + dv) * R + e] : 0.f;
    float m = -INFINITY, s = 0.f;
    const int t0 = blockIdx.y * tchunk, t1 = min(T, t0 + tchunk);
    for (int tb = t0; tb < t1; tb += SC_TT) {
        const int nt = min(SC_TT, t1 - tb);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * R; k += 128)  // nt * 4 * R floats, float4 at a time
            reinterpret_cast<float4 *>(tile)[k] = reinterpret_cast<const float4 *>(x2 + (size_t)tb * 4 * R)[k];
        __syncthreads();
        int t = 0;
        for (; t + 4 <= nt; t += 4) {
            float sc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float *p = tile + ((t + u) * 4 + dv) * R;
                float acc = 0.f;
#pragma unroll
                for (int e = 0; e < R; ++e) acc = fmaf(a[e], p[e], acc);
                sc[u] = acc;
            }
            const float mx = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
            s = s * ex2f((m - mx) * LOG2E_F);
#pragma unroll
            for (int u = 0; u < 4; ++u) s += ex2f((sc[u] - mx) * LOG2E_F);
            m = mx;
        }
        for (; t < nt; ++t) {
            const float *p = tile + (t * 4 + dv) * R;
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) acc = fmaf(a[e], p[e], acc);
            lse_merge(m, s, acc, 1.f);
        }
    }
    if (live) part[((size_t)blockIdx.y * npair + pair) * 4 + dv] = make_float2(m, s);
}

// log-sum-exp of the root scores over the vocabulary (one CTA)
__global__ void __launch_bounds__(1024) root_lse_kernel(const float *__restrict__ root_score, int T, float *__restrict__ out) {
    __shared__ float sm[32], ss[32];
    float m = -INFINITY, s = 0.f;
    for (int t = threadIdx.x; t < T; t += 1024) lse_merge(m, s, root_score[t], 1.f);
    for (int o = 16; o >= 1; o >>= 1) lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x < 32) {
        m = sm[threadIdx.x]; s = ss[threadIdx.x];
        for (int o = 16; o >= 1; o >>= 1) lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
        if (threadIdx.x == 0) out[0] = m + __logf(s);
    }
}

// ---- pass 2: the merged tensors ----------------------------------------------------------------------------------
// one CTA per (b, merged head row hp = 0..n); threads over the (child, valence) entries of the row.
template <int R>
__global__ void __launch_bounds__(128) scores_write_kernel(const float *__restrict__ x1, const float *__restrict__ x2,
                                                           const long long *__restrict__ token, const unsigned char *__restrict__ head_mask,
                                                           const float *__restrict__ dec_score, const float *__restrict__ root_score,
                                                           const float2 *__restrict__ part, int tsplit, const float *__restrict__ root_lse,
                                                           int B, int n, float one, float zero, float neg_fill,
                                                           float *__restrict__ mdec, float *__restrict__ mattach, float *__restrict__ lse_out) {
    __shared__ float s_x1[4 * R];
    __shared__ float s_lse[4];
    const int N = n + 1;
    const int b = blockIdx.x / N, hp = blockIdx.x % N;
    float *row = mattach + ((size_t)b * N + hp) * N * 2;
    float *drow = mdec + ((size_t)b * N + hp) * 8;
    if (hp == 0) {  // ROOT: distributions.py:256-257, 262
        const float rl = root_lse[0];
        for (int k = threadIdx.x; k < N * 2; k += 128) {
            const int cp = k >> 1, v = k & 1;
            float x = zero;
            if (cp >= 1 && v == 1) x = root_score[token[(size_t)b * n + cp - 1]] - rl;
            row[k] = x;
        }
        if (threadIdx.x < 8) drow[threadIdx.x] = (threadIdx.x >> 2) == 1 ? one : zero;
        return;
    }
    const int h = hp - 1;
    const size_t pair = (size_t)b * n + h;
    const int npair = B * n;
    for (int k = threadIdx.x; k < 4 * R; k += 128) s_x1[k] = x1[pair * 4 * R + k];
    if (threadIdx.x < 4) {
        float m = -INFINITY, s = 0.f;
        for (int y = 0; y < tsplit; ++y) {
            const float2 p = part[((size_t)y * npair + pair) * 4 + threadIdx.x];
            lse_merge(m, s, p.x, p.y);
        }
        const float l = m + __logf(s);
        s_lse[threadIdx.x] = l;
        lse_out[pair * 4 + threadIdx.x] = l;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 40) {  // dec[b,h,d,v,k] = log_softmax_k(dec_score[b,h,k,d,v])   (ldndmv.py:202)
        const int q = threadIdx.x - 32, dvi = q >> 1, k = q & 1;
        const float s0 = dec_score[pair * 8 + dvi], s1 = dec_score[pair * 8 + 4 + dvi];
        const float mx = fmaxf(s0, s1);
        const float l = mx + __logf(__expf(s0 - mx) + __expf(s1 - mx));
        drow[q] = (k ? s1 : s0) - l;
    }
    __syncthreads();
    const bool masked = head_mask && head_mask[pair];
    for (int k = threadIdx.x; k < N * 2; k += 128) {
        const int cp = k >> 1, v = k & 1;
        float x = zero;
        if (cp >= 1) {
            const int c = cp - 1;
            if (masked) x = neg_fill;          // masked_fill_ over the whole head row, diagonal included (:194-198)
            else if (c == h) x = 0.f;          // both triangular masks are 0 on the diagonal (:190-193)
            else {
                const int dvi = (c < h ? 0 : 2) + v;
                const float *p = x2 + ((size_t)token[(size_t)b * n + c] * 4 + dvi) * R;
                float acc = 0.f;
#pragma unroll
                for (int e = 0; e < R; ++e) acc = fmaf(s_x1[dvi * R + e], p[e], acc);
                x = acc - s_lse[dvi];
            }
        }
        row[k] = x;
    }
}

// ---- backward, direct terms ----------------------------------------------------------------------------------------
// one CTA per (b, hp).  d x1 += sum_c g p2[token_c];  d x2[token_c] += g p1;  G[b,h,d,v] = sum_c g;  dec / root terms.
template <int R>
__global__ void __launch_bounds__(128) scores_bwd_direct_kernel(const float *__restrict__ x1, const float *__restrict__ x2,
                                                                const long long *__restrict__ token, const unsigned char *__restrict__ head_mask,
                                                                const float *__restrict__ dec_score, const float *__restrict__ g_mdec,
                                                                const float *__restrict__ g_mattach, int B, int n, float *__restrict__ g_x1,
                                                                float *__restrict__ g_x2, float *__restrict__ g_dec_score,
                                                                float *__restrict__ g_root_score, float *__restrict__ G, float *__restrict__ g_root_total) {
    __shared__ float s_x1[4 * R];
    __shared__ float s_acc[4 * R];
    __shared__ float s_G[4];
    const int N = n + 1;
    const int b = blockIdx.x / N, hp = blockIdx.x % N;
    const float *grow = g_mattach + ((size_t)b * N + hp) * N * 2;
    if (hp == 0) {
        float tot = 0.f;
        for (int cp = 1 + threadIdx.x; cp < N; cp += 128) {
            const float g = grow[cp * 2 + 1];
            if (g != 0.f) { atomicAdd(&g_root_score[token[(size_t)b * n + cp - 1]], g); tot += g; }
        }
        for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if ((threadIdx.x & 31) == 0 && tot != 0.f) atomicAdd(g_root_total, tot);
        return;
    }
    const int h = hp - 1;
    const size_t pair = (size_t)b * n + h;
    for (int k = threadIdx.x; k < 4 * R; k += 128) { s_x1[k] = x1[pair * 4 * R + k]; s_acc[k] = 0.f; }
    if (threadIdx.x < 4) s_G[threadIdx.x] = 0.f;
    if (threadIdx.x >= 32 && threadIdx.x < 36) {  // log_softmax backward over the two decisions
        const int dvi = threadIdx.x - 32;
        const float s0 = dec_score[pair * 8 + dvi], s1 = dec_score[pair * 8 + 4 + dvi];
        const float mx = fmaxf(s0, s1);
        const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
        const float g0 = g_mdec[((size_t)b * N + hp) * 8 + dvi * 2], g1 = g_mdec[((size_t)b * N + hp) * 8 + dvi * 2 + 1];
        const float gs = g0 + g1, inv = 1.f / (e0 + e1);
        g_dec_score[pair * 8 + dvi] = g0 - gs * e0 * inv;
        g_dec_score[pair * 8 + 4 + dvi] = g1 - gs * e1 * inv;
    }
    __syncthreads();
    const bool masked = head_mask && head_mask[pair];
    if (!masked) {
        // thread = (child slot, e-quarter): 4 threads share a (c, v) entry? keep it simple: one thread per (c, v) entry
        for (int k = threadIdx.x; k < n * 2; k += 128) {
            const int c = k >> 1, v = k & 1;
            if (c == h) continue;
            const float g = grow[(c + 1) * 2 + v];
            if (g == 0.f) continue;
            const int dvi = (c < h ? 0 : 2) + v;
            const size_t tok = (size_t)token[(size_t)b * n + c];
            const float *p = x2 + (tok * 4 + dvi) * R;
            float *gp = g_x2 + (tok * 4 + dvi) * R;
#pragma unroll
            for (int e = 0; e < R; ++e) {
                atomicAdd(&s_acc[dvi * R + e], g * p[e]);
                atomicAdd(&gp[e], g * s_x1[dvi * R + e]);
            }
            atomicAdd(&s_G[dvi], g);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * R; k += 128)
        if (s_acc[k] != 0.f) atomicAdd(&g_x1[pair * 4 * R + k], s_acc[k]);
    if (threadIdx.x < 4) G[pair * 4 + threadIdx.x] = s_G[threadIdx.x];
}

// ---- backward, softmax term of d x1: d x1[row] -= G[row] sum_t softmax[row][t] x2[t] -------------------------------
template <int R>
__global__ void __launch_bounds__(128) scores_bwd_rows_kernel(const float *__restrict__ x1, const float *__restrict__ x2,
                                                              const float *__restrict__ lse, const float *__restrict__ G, int npair,
                                                              int T, int tchunk, float *__restrict__ g_x1) {
    __shared__ __align__(16) float tile[SC_TT * 4 * R];
    const int dv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 32 + lane;
    const bool live = pair < npair;
    float a[R], acc[R];
#pragma unroll
    for (int e = 0; e < R; ++e) { a[e] = live ? x1[((size_t)pair * 4 + dv) * R + e] : 0.f; acc[e] = 0.f; }
    const float g = live ? G[(size_t)pair * 4 + dv] : 0.f;
    const float l2 = live ? lse[(size_t)pair * 4 + dv] * LOG2E_F : 0.f;
    if (__syncthreads_or(g != 0.f) == 0) return;  // no gradient reaches any row of this CTA (padding heads)
    const int t0 = blockIdx.y * tchunk, t1 = min(T, t0 + tchunk);
    for (int tb = t0; tb < t1; tb += SC_TT) {
        const int nt = min(SC_TT, t1 - tb);
        __syncthreads();
        for (int k = threadIdx.x; k < nt * R; k += 128)
            reinterpret_cast<float4 *>(tile)[k] = reinterpret_cast<const float4 *>(x2 + (size_t)tb * 4 * R)[k];
        __syncthreads();
        for (int t = 0; t < nt; ++t) {
            const float *p = tile + (t * 4 + dv) * R;
            float sc = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) sc = fmaf(a[e], p[e], sc);
            const float w = ex2f(fmaf(sc, LOG2E_F, -l2));
#pragma unroll
            for (int e = 0; e < R; ++e) acc[e] = fmaf(w, p[e], acc[e]);
        }
    }
    if (live && g != 0.f) {
#pragma unroll
        for (int e = 0; e < R; ++e) atomicAdd(&g_x1[((size_t)pair * 4 + dv) * R + e], -g * acc[e]);
    }
}

// ---- backward, softmax term of d x2: d x2[t] -= sum_rows G[row] softmax[row][t] x1[row] ----------------------------
// grid (ceil(T / 32), rsplit); warp = (d, v), lane = token; the (b,h) pairs stream through shared memory.
template <int R>
__global__ void __launch_bounds__(128) scores_bwd_tokens_kernel(const float *__restrict__ x1, const float *__restrict__ x2,
                                                                const float *__restrict__ lse, const float *__restrict__ G, int npair,
                                                                int T, int pchunk, float *__restrict__ g_x2) {
    __shared__ __align__(16) float tile[SC_TT * 4 * R];
    __shared__ float s_l[SC_TT * 4], s_g[SC_TT * 4];
    const int dv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 32 + lane;
    const bool live = t < T;
    float a[R], acc[R];
#pragma unroll
    for (int e = 0; e < R; ++e) { a[e] = live ? x2[((size_t)t * 4 + dv) * R + e] : 0.f; acc[e] = 0.f; }
    const int p0 = blockIdx.y * pchunk, p1 = min(npair, p0 + pchunk);
    for (int pb = p0; pb < p1; pb += SC_TT) {
        const int np = min(SC_TT, p1 - pb);
        __syncthreads();
        for (int k = threadIdx.x; k < np * R; k += 128)
            reinterpret_cast<float4 *>(tile)[k] = reinterpret_cast<const float4 *>(x1 + (size_t)pb * 4 * R)[k];
        for (int k = threadIdx.x; k < np * 4; k += 128) { s_l[k] = lse[(size_t)pb * 4 + k] * LOG2E_F; s_g[k] = G[(size_t)pb * 4 + k]; }
        __syncthreads();
        for (int q = 0; q < np; ++q) {
            const float g = s_g[q * 4 + dv];
            if (g == 0.f) continue;  // warp-uniform
            const float *p = tile + (q * 4 + dv) * R;
            float sc = 0.f;
#pragma unroll
            for (int e = 0; e < R; ++e) sc = fmaf(a[e], p[e], sc);
            const float w = g * ex2f(fmaf(sc, LOG2E_F, -s_l[q * 4 + dv]));
#pragma unroll
            for (int e = 0; e < R; ++e) acc[e] = fmaf(w, p[e], acc[e]);
        }
    }
    if (live) {
#pragma unroll
        for (int e = 0; e < R; ++e)
            if (acc[e] != 0.f) atomicAdd(&g_x2[((size_t)t * 4 + dv) * R + e], -acc[e]);
    }
}

// d root_score[t] -= softmax(root_score)[t] * (sum of the root gradients)
__global__ void root_bwd_kernel(const float *__restrict__ root_score, const float *__restrict__ root_lse,
                                const float *__restrict__ g_total, int T, float *__restrict__ g_root_score) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) g_root_score[t] -= __expf(root_score[t] - root_lse[0]) * g_total[0];
}

int split_for(int units, int per_unit_ctas, int sm_count) {
    // enough CTAs for ~4 per SM, at most 16 slices
    int s = (4 * sm_count + per_unit_ctas - 1) / (per_unit_ctas > 0 ? per_unit_ctas : 1);
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    if (s > units) s = units > 0 ? units : 1;
    return s;
}

}  // namespace

// workspace: partial (m, s) of the row pass [16][B n][4] float2 | G [B n][4] | root gradient total [1] (+ padding)
size_t dmv_scores_workspace_bytes(int B, int n) {
    const size_t npair = (size_t)B * n;
    return 16 * npair * 4 * sizeof(float2) + npair * 4 * sizeof(float) + 256;
}

template <int R>
static cudaError_t scores_forward_r(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                                    const float *dec_score, const float *root_score, int B, int n, int T, float one, float zero,
                                    float neg_fill, float *mdec, float *mattach, float *lse, float *root_lse, void *ws,
                                    int sm_count, cudaStream_t st) {
    const int npair = B * n, gx = (npair + 31) / 32;
    float2 *part = reinterpret_cast<float2 *>(ws);
    int tsplit = split_for((T + SC_TT - 1) / SC_TT, gx, sm_count);
    int tchunk = ((T + tsplit - 1) / tsplit + SC_TT - 1) / SC_TT * SC_TT;
    tsplit = (T + tchunk - 1) / tchunk;
    scores_lse_kernel<R><<<dim3(gx, tsplit), 128, 0, st>>>(x1, x2, npair, T, tchunk, part);
    root_lse_kernel<<<1, 1024, 0, st>>>(root_score, T, root_lse);
    scores_write_kernel<R><<<B * (n + 1), 128, 0, st>>>(x1, x2, token, head_mask, dec_score, root_score, part, tsplit, root_lse, B, n,
                                                       one, zero, neg_fill, mdec, mattach, lse);
    return cudaGetLastError();
}

cudaError_t launch_dmv_scores(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                              const float *dec_score, const float *root_score, int B, int n, int T, int r, float one, float zero,
                              float neg_fill, float *mdec, float *mattach, float *lse, float *root_lse, void *ws, int sm_count,
                              cudaStream_t st) {
    switch (r) {
        case 4: return scores_forward_r<4>(x1, x2, token, head_mask, dec_score, root_score, B, n, T, one, zero, neg_fill, mdec, mattach, lse, root_lse, ws, sm_count, st);
        case 8: return scores_forward_r<8>(x1, x2, token, head_mask, dec_score, root_score, B, n, T, one, zero, neg_fill, mdec, mattach, lse, root_lse, ws, sm_count, st);
        case 16: return scores_forward_r<16>(x1, x2, token, head_mask, dec_score, root_score, B, n, T, one, zero, neg_fill, mdec, mattach, lse, root_lse, ws, sm_count, st);
        case 32: return scores_forward_r<32>(x1, x2, token, head_mask, dec_score, root_score, B, n, T, one, zero, neg_fill, mdec, mattach, lse, root_lse, ws, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int R>
static cudaError_t scores_backward_r(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                                     const float *dec_score, const float *root_score, const float *lse, const float *root_lse,
                                     const float *g_mdec, const float *g_mattach, int B, int n, int T, float *g_x1, float *g_x2,
                                     float *g_dec_score, float *g_root_score, void *ws, int sm_count, cudaStream_t st) {
    const int npair = B * n;
    float *G = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + 16 * (size_t)npair * 4 * sizeof(float2));
    float *g_total = G + (size_t)npair * 4;
    cudaError_t e;
    if ((e = cudaMemsetAsync(g_x1, 0, (size_t)npair * 4 * R * sizeof(float), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g_x2, 0, (size_t)T * 4 * R * sizeof(float), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g_root_score, 0, (size_t)T * sizeof(float), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(g_total, 0, sizeof(float), st)) != cudaSuccess) return e;
    scores_bwd_direct_kernel<R><<<B * (n + 1), 128, 0, st>>>(x1, x2, token, head_mask, dec_score, g_mdec, g_mattach, B, n, g_x1, g_x2,
                                                            g_dec_score, g_root_score, G, g_total);
    const int gx = (npair + 31) / 32;
    int tsplit = split_for((T + SC_TT - 1) / SC_TT, gx, sm_count);
    int tchunk = ((T + tsplit - 1) / tsplit + SC_TT - 1) / SC_TT * SC_TT;
    tsplit = (T + tchunk - 1) / tchunk;
    scores_bwd_rows_kernel<R><<<dim3(gx, tsplit), 128, 0, st>>>(x1, x2, lse, G, npair, T, tchunk, g_x1);
    const int gt = (T + 31) / 32;
    int psplit = split_for((npair + SC_TT - 1) / SC_TT, gt, sm_count);
    int pchunk = ((npair + psplit - 1) / psplit + SC_TT - 1) / SC_TT * SC_TT;
    psplit = (npair + pchunk - 1) / pchunk;
    scores_bwd_tokens_kernel<R><<<dim3(gt, psplit), 128, 0, st>>>(x1, x2, lse, G, npair, T, pchunk, g_x2);
    root_bwd_kernel<<<(T + 255) / 256, 256, 0, st>>>(root_score, root_lse, g_total, T, g_root_score);
    return cudaGetLastError();
}

cudaError_t launch_dmv_scores_backward(const float *x1, const float *x2, const long long *token, const unsigned char *head_mask,
                                       const float *dec_score, const float *root_score, const float *lse, const float *root_lse,
                                       const float *g_mdec, const float *g_mattach, int B, int n, int T, int r, float *g_x1,
                                       float *g_x2, float *g_dec_score, float *g_root_score, void *ws, int sm_count, cudaStream_t st) {
    switch (r) {
        case 4: return scores_backward_r<4>(x1, x2, token, head_mask, dec_score, root_score, lse, root_lse, g_mdec, g_mattach, B, n, T, g_x1, g_x2, g_dec_score, g_root_score, ws, sm_count, st);
        case 8: return scores_backward_r<8>(x1, x2, token, head_mask, dec_score, root_score, lse, root_lse, g_mdec, g_mattach, B, n, T, g_x1, g_x2, g_dec_score, g_root_score, ws, sm_count, st);
        case 16: return scores_backward_r<16>(x1, x2, token, head_mask, dec_score, root_score, lse, root_lse, g_mdec, g_mattach, B, n, T, g_x1, g_x2, g_dec_score, g_root_score, ws, sm_count, st);
        case 32: return scores_backward_r<32>(x1, x2, token, head_mask, dec_score, root_score, lse, root_lse, g_mdec, g_mattach, B, n, T, g_x1, g_x2, g_dec_score, g_root_score, ws, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vlgae
