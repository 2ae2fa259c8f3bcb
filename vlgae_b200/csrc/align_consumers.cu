// align_consumers.cu -- the small fp32 kernels behind the fused grounding consumers.
//
// The reference materialises attmap [B][A][Q][V] (7.4 GB at the cfg2 shape) and then only ever looks at
//   max over V  -> log_softmax over A -> diagonal a = b                       loss txt2vis   joint.py:473-478
//   max over Q  -> log_softmax over B -> diagonal                             loss vis2txt   joint.py:480-483
//   max over V, max over A -> indices                                         decode         joint.py:520
//   the diagonal slab attmap[b, b] [Q][V] (+ POS priors, heuristics, top 5)   decode / loss  joint.py:466-469, 522-594
// The two maxima come out of the tcgen05 kernel's epilogue (align_kernels.cu, vlgae_align_maxima); this file holds
//   * align_diagonal_kernel      the diagonal slab as a plain fp32 contraction (B small GEMMs of Q x V x D)
//   * grounding_ce_kernel        both cross-entropy sums from the two maxima (log-softmax over A / B, diagonal, weights)
//   * topk_rows_kernel           the 5 best factors of every (caption, query) row (argsort(-1, descending)[..., :5])
//   * max_backward_kernel        backward of max over V: routes g[b,a,q] to txt[b,q,:] and vis[a, argv, :]
#include <cuda_runtime.h>
#include <stdint.h>

#include "align_kernels.cuh"

namespace vlgae {
namespace {

// ---------------------------------------------------------------------------------------------
// out[b][q][v] = <txt[b,q,:], vis[b,v,:]>, neg where a mask is off.  CTA = (v-tile of 64, caption); the vis tile is held
// transposed ([d][v]: conflict-free across the lanes' factors), the caption's queries row-major (broadcast reads).
// ---------------------------------------------------------------------------------------------
constexpr int DV = 64;   // factors per CTA
constexpr int DQ = 4;    // query groups (threads = DV * DQ)
__global__ void __launch_bounds__(DV * DQ) align_diagonal_kernel(const float *__restrict__ vis, const uint8_t *__restrict__ vis_mask,
                                                                 const float *__restrict__ txt, const uint8_t *__restrict__ txt_mask,
                                                                 int V, int Q, int D, float neg, float *__restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.y, v0 = blockIdx.x * DV;
    const int Dp = (D + 3) & ~3;
    float *vs = sm;                  // [Dp][DV]
    float *ts = sm + (size_t)Dp * DV;  // [Q][Dp]
    for (int t = threadIdx.x; t < DV * Dp; t += DV * DQ) {
        const int v = t / Dp, d = t - v * Dp;
        vs[d * DV + v] = (v0 + v < V && d < D) ? vis[((size_t)b * V + v0 + v) * D + d] : 0.f;
    }
    for (int t = threadIdx.x; t < Q * Dp; t += DV * DQ) {
        const int q = t / Dp, d = t - q * Dp;
        ts[t] = d < D ? txt[((size_t)b * Q + q) * D + d] : 0.f;
    }
    __syncthreads();
    const int v = threadIdx.x % DV, qg = threadIdx.x / DV;
    const bool v_ok = v0 + v < V;
    const bool v_keep = v_ok && vis_mask[(size_t)b * V + v0 + v] != 0;
    constexpr int QB = 8;  // queries per register block
    for (int q0 = qg * QB; q0 < Q; q0 += DQ * QB) {
        float acc[QB];
#pragma unroll
        for (int k = 0; k < QB; ++k) acc[k] = 0.f;
        for (int d = 0; d < Dp; d += 4) {
            const float x0 = vs[(d + 0) * DV + v], x1 = vs[(d + 1) * DV + v], x2 = vs[(d + 2) * DV + v], x3 = vs[(d + 3) * DV + v];
#pragma unroll
            for (int k = 0; k < QB; ++k) {
                if (q0 + k < Q) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(ts + (size_t)(q0 + k) * Dp + d);
                    acc[k] = fmaf(x0, t4.x, fmaf(x1, t4.y, fmaf(x2, t4.z, fmaf(x3, t4.w, acc[k]))));
                }
            }
        }
        if (v_ok) {
#pragma unroll
            for (int k = 0; k < QB; ++k)
                if (q0 + k < Q) {
                    const bool keep = v_keep && txt_mask[(size_t)b * Q + q0 + k] != 0;
                    out[((size_t)b * Q + q0 + k) * V + v0 + v] = keep ? acc[k] : neg;
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// out2[0] = -sum_{b,q} log_softmax_a(maxv[b,:,q])[b] * txt_marginal[b,q]      (joint.py:473-478)
// out2[1] = -sum_{a,v} log_softmax_b(maxq[:,a,v])[a] * vis_mask[a,v]          (joint.py:480-483)
// one thread per (b, q) resp. (a, v): two passes over the 128 entries of its softmax axis (max, then sum of exp)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float x, float *red) {
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;
}

__global__ void grounding_ce_kernel(const float *__restrict__ maxv, const float *__restrict__ maxq, const float *__restrict__ marg,
                                    const uint8_t *__restrict__ vis_mask, int B, int Q, int V, float *out2) {
    __shared__ float red[32];
    const long long n1 = (long long)B * Q, n2 = maxq ? (long long)B * V : 0;
    float s1 = 0.f, s2 = 0.f;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n1 + n2; t += (long long)gridDim.x * blockDim.x) {
        if (t < n1) {
            const int b = (int)(t / Q), q = (int)(t - (long long)b * Q);
            const float w = marg[t];
            const float *p = maxv + (size_t)b * B * Q + q;  // [b][a][q], a strides Q
            float m = p[0];
            for (int a = 1; a < B; ++a) m = fmaxf(m, p[(size_t)a * Q]);
            float s = 0.f;
            for (int a = 0; a < B; ++a) s += __expf(p[(size_t)a * Q] - m);
            s1 -= (p[(size_t)b * Q] - m - __logf(s)) * w;
        } else {
            const long long u = t - n1;
            const int a = (int)(u / V), v = (int)(u - (long long)a * V);
            const float w = vis_mask[u] ? 1.f : 0.f;
            const float *p = maxq + (size_t)a * V + v;  // [b][a][v], b strides B * V
            const size_t st = (size_t)B * V;
            float m = p[0];
            for (int bb = 1; bb < B; ++bb) m = fmaxf(m, p[bb * st]);
            float s = 0.f;
            for (int bb = 0; bb < B; ++bb) s += __expf(p[bb * st] - m);
            s2 -= (p[a * st] - m - __logf(s)) * w;
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { atomicAdd(out2, s1); atomicAdd(out2 + 1, s2); }
}

// ---------------------------------------------------------------------------------------------
// idx[row][0..k) = the k largest entries of x[row][0..V) in descending order, smaller index first on ties (k <= 8).
// One warp per row: every lane keeps the sorted top k of its strided slice, then k rounds of a warp-wide arg-max.
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_MAX = 8;
__global__ void topk_rows_kernel(const float *__restrict__ x, long long rows, int V, int k, int *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = wid; row < rows; row += nw) {
        const float *xr = x + row * V;
        float bv[TOPK_MAX];
        int bi[TOPK_MAX];
#pragma unroll
        for (int j = 0; j < TOPK_MAX; ++j) { bv[j] = -INFINITY; bi[j] = 0x7fffffff; }
        for (int v = lane; v < V; v += 32) {
            float cv = xr[v];
            int ci = v;
            if (cv > bv[TOPK_MAX - 1] || (cv == bv[TOPK_MAX - 1] && ci < bi[TOPK_MAX - 1])) {
#pragma unroll
                for (int j = 0; j < TOPK_MAX; ++j) {  // insertion into the sorted list (bubble the displaced entry down)
                    const bool better = cv > bv[j] || (cv == bv[j] && ci < bi[j]);
                    const float tv = bv[j];
                    const int ti = bi[j];
                    if (better) { bv[j] = cv; bi[j] = ci; cv = tv; ci = ti; }
                }
            }
        }
        int head = 0;  // this lane's best not yet emitted
        for (int r = 0; r < k; ++r) {
            float cv = -INFINITY;
            int ci = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < TOPK_MAX; ++j)
                if (j == head) { cv = bv[j]; ci = bi[j]; }
            float mv = cv;
            int mi = ci;
            for (int o = 16; o >= 1; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, mv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (ov > mv || (ov == mv && oi < mi)) { mv = ov; mi = oi; }
            }
            if (ci == mi && head < TOPK_MAX) ++head;
            if (lane == 0) idx[row * k + r] = mi == 0x7fffffff ? 0 : mi;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward of maxv[b,a,q] = max_v att[b,a,q,v] with att = <txt[b,q,:], vis[a,v,:]> and the masks as constants:
//   v* = argv[b,a,q];  if txt_mask[b,q] and vis_mask[a,v*]:  d txt[b,q,:] += g vis[a,v*,:],  d vis[a,v*,:] += g txt[b,q,:]
// one warp per (b, q): the caption's gradient row is accumulated in registers over a (no atomics); the image side is
// scattered with atomics (B * A * Q rows of D floats).
// ---------------------------------------------------------------------------------------------
__global__ void max_backward_kernel(const float *__restrict__ g, const int *__restrict__ argv, const float *__restrict__ vis,
                                    const uint8_t *__restrict__ vis_mask, const float *__restrict__ txt,
                                    const uint8_t *__restrict__ txt_mask, int A, int V, int B, int Q, int D, float *grad_vis,
                                    float *grad_txt) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long bq = wid; bq < (long long)B * Q; bq += nw) {
        const int b = (int)(bq / Q), q = (int)(bq - (long long)b * Q);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};  // D <= 128: 4 floats per lane
        float tq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tq[k] = lane + 32 * k < D ? txt[bq * D + lane + 32 * k] : 0.f;
        if (txt_mask[bq]) {
            for (int a = 0; a < A; ++a) {
                const size_t o = ((size_t)b * A + a) * Q + q;
                const float gg = g[o];
                const int v = argv[o];
                if (gg == 0.f || !vis_mask[(size_t)a * V + v]) continue;
                const float *vr = vis + ((size_t)a * V + v) * D;
                float *gv = grad_vis ? grad_vis + ((size_t)a * V + v) * D : nullptr;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int d = lane + 32 * k;
                    if (d < D) {
                        acc[k] = fmaf(gg, vr[d], acc[k]);
                        if (gv) atomicAdd(gv + d, gg * tq[k]);
                    }
                }
            }
        }
        if (grad_txt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < D) grad_txt[bq * D + lane + 32 * k] = acc[k];
        }
    }
}

}  // namespace

cudaError_t launch_align_diagonal(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int B,
                                  int V, int Q, int D, float neg, float *out, cudaStream_t st) {
    const int Dp = (D + 3) & ~3;
    const size_t smem = ((size_t)Dp * DV + (size_t)Q * Dp) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(align_diagonal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((V + DV - 1) / DV, B);
    align_diagonal_kernel<<<grid, DV * DQ, smem, st>>>(vis, vis_mask, txt, txt_mask, V, Q, D, neg, out);
    return cudaGetLastError();
}

cudaError_t launch_grounding_ce(const float *maxv, const float *maxq, const float *txt_marginal, const uint8_t *vis_mask,
                                int B, int Q, int V, float *out2, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(float), st);
    if (e != cudaSuccess) return e;
    const long long n = (long long)B * Q + (maxq ? (long long)B * V : 0);
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    grounding_ce_kernel<<<grid, 256, 0, st>>>(maxv, maxq, txt_marginal, vis_mask, B, Q, V, out2);
    return cudaGetLastError();
}

cudaError_t launch_topk_rows(const float *x, long long rows, int V, int k, int *idx, cudaStream_t st) {
    if (k < 1 || k > TOPK_MAX) return cudaErrorInvalidValue;
    long long warps = rows;
    int grid = (int)((warps * 32 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    topk_rows_kernel<<<grid, 256, 0, st>>>(x, rows, V, k, idx);
    return cudaGetLastError();
}

cudaError_t launch_max_over_factors_backward(const float *g, const int *argv, const float *vis, const uint8_t *vis_mask,
                                             const float *txt, const uint8_t *txt_mask, int A, int V, int B, int Q, int D,
                                             float *grad_vis, float *grad_txt, cudaStream_t st) {
    if (D > 128) return cudaErrorInvalidValue;
    if (grad_vis) {
        cudaError_t e = cudaMemsetAsync(grad_vis, 0, (size_t)A * V * D * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    const long long warps = (long long)B * Q;
    int grid = (int)((warps * 32 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    max_backward_kernel<<<grid, 256, 0, st>>>(g, argv, vis, vis_mask, txt, txt_mask, A, V, B, Q, D, grad_vis, grad_txt);
    return cudaGetLastError();
}

}  // namespace vlgae
