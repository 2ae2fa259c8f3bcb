// word_attention.cu -- per-sample word -> factor attention of DependencyBoxRel._forward.
//
// Reference (/root/reference/src/model/joint.py:668-673):
//     attmap = einsum("bvd,bqd->bqv", vis, txt[:, 1:]).softmax(2);  x = einsum("bqv,bvh->bqh", attmap, vis_mid)
// i.e. one small attention per caption: n <= 64 word queries over V factors (no mask), values vis_mid.  A fraction of a
// per cent of the alignment contraction's flops, so a plain fp32 kernel: one CTA per caption, the caption's queries
// resident in shared memory, the factors streamed in tiles of 32 with an online softmax; thread h owns output column h
// for every query (registers).  The backward recomputes the probabilities from the saved row log-sum-exp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "align_kernels.cuh"

namespace vlgae {
namespace {

constexpr int WA_T = 256;   // threads (= max H)
constexpr int WA_TV = 32;   // factors per tile
constexpr int WA_NQ = 64;   // max queries

// shared: txt [n][D] | vis tile [TV][D] | mid tile [TV][H] | P [n][TV] | m [n] | l [n] | alpha [n]
// (tile rows are padded by one float: lanes that walk the factors of a tile then hit distinct banks)
template <bool BWD>
__host__ __device__ inline size_t wa_smem(int n, int D, int H) {
    size_t f = (size_t)n * D + (size_t)WA_TV * (D + 1) + (size_t)WA_TV * (H + 1) + (size_t)n * WA_TV + 3 * (size_t)n;
    if (BWD) f += (size_t)n * H + (size_t)n * WA_TV + (size_t)n * D;  // dO [n][H] | dS [n][TV] | dtxt [n][D]
    return f * sizeof(float);
}

__global__ void __launch_bounds__(WA_T) word_attn_fwd_kernel(const float *__restrict__ vis, const float *__restrict__ txt,
                                                             const float *__restrict__ mid, int V, int n, int D, int H,
                                                             float *__restrict__ out, float *__restrict__ lse) {
    extern __shared__ __align__(16) float sm[];
    const int DP = D + 1, HP = H + 1;
    float *s_txt = sm, *s_vis = s_txt + n * D, *s_mid = s_vis + WA_TV * DP, *s_p = s_mid + WA_TV * HP;
    float *s_m = s_p + n * WA_TV, *s_l = s_m + n, *s_a = s_l + n;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *vb = vis + (size_t)b * V * D, *mb = mid + (size_t)b * V * H;
    for (int t = tid; t < n * D; t += WA_T) s_txt[t] = txt[(size_t)b * n * D + t];
    for (int q = tid; q < n; q += WA_T) { s_m[q] = -INFINITY; s_l[q] = 0.f; }
    float acc[WA_NQ];
#pragma unroll
    for (int q = 0; q < WA_NQ; ++q) acc[q] = 0.f;
    __syncthreads();
    for (int v0 = 0; v0 < V; v0 += WA_TV) {
        const int tv = min(WA_TV, V - v0);
        for (int t = tid; t < tv * D; t += WA_T) s_vis[(t / D) * DP + t % D] = vb[(size_t)v0 * D + t];
        for (int t = tid; t < tv * H; t += WA_T) s_mid[(t / H) * HP + t % H] = mb[(size_t)v0 * H + t];
        __syncthreads();
        // scores S[q][v] = <txt[q], vis[v]>
        for (int e = tid; e < n * WA_TV; e += WA_T) {
            const int q = e / WA_TV, v = e - q * WA_TV;
            float s = -INFINITY;
            if (v < tv) {
                s = 0.f;
                const float *a = s_txt + q * D, *c = s_vis + v * DP;
                for (int d = 0; d < D; ++d) s = fmaf(a[d], c[d], s);
            }
            s_p[e] = s;
        }
        __syncthreads();
        // online softmax per query row (one thread per row: 32 entries)
        if (tid < n) {
            float *row = s_p + tid * WA_TV;
            float mx = s_m[tid];
            for (int v = 0; v < tv; ++v) mx = fmaxf(mx, row[v]);
            const float al = __expf(s_m[tid] - mx);
            float sum = 0.f;
            for (int v = 0; v < WA_TV; ++v) {
                const float pv = v < tv ? __expf(row[v] - mx) : 0.f;
                row[v] = pv;
                sum += pv;
            }
            s_l[tid] = s_l[tid] * al + sum;
            s_m[tid] = mx;
            s_a[tid] = al;
        }
        __syncthreads();
        if (tid < H) {
#pragma unroll  // fully unrolled: acc[] stays in registers
            for (int q = 0; q < WA_NQ; ++q) {
                if (q < n) {
                    float o = acc[q] * s_a[q];
                    const float *pr = s_p + q * WA_TV;
                    for (int v = 0; v < tv; ++v) o = fmaf(pr[v], s_mid[v * HP + tid], o);
                    acc[q] = o;
                }
            }
        }
        __syncthreads();
    }
    if (tid < H) {
#pragma unroll
        for (int q = 0; q < WA_NQ; ++q)
            if (q < n) out[((size_t)b * n + q) * H + tid] = acc[q] / s_l[q];
    }
    if (lse && tid < n) lse[(size_t)b * n + tid] = s_m[tid] + __logf(s_l[tid]);
}

// backward: P = exp(S - lse);  dmid = P^T dO;  dP = dO mid^T;  dS = P (dP - rowsum(dO * O));  dvis = dS^T txt;  dtxt = dS vis
__global__ void __launch_bounds__(WA_T) word_attn_bwd_kernel(const float *__restrict__ vis, const float *__restrict__ txt,
                                                             const float *__restrict__ mid, const float *__restrict__ out,
                                                             const float *__restrict__ lse, const float *__restrict__ gout,
                                                             int V, int n, int D, int H, float *__restrict__ gvis,
                                                             float *__restrict__ gtxt, float *__restrict__ gmid) {
    extern __shared__ __align__(16) float sm[];
    const int DP = D + 1, HP = H + 1;
    float *s_txt = sm, *s_vis = s_txt + n * D, *s_mid = s_vis + WA_TV * DP, *s_p = s_mid + WA_TV * HP;
    float *s_lse = s_p + n * WA_TV, *s_dr = s_lse + n, *s_unused = s_dr + n;
    float *s_do = s_unused + n, *s_ds = s_do + n * H, *s_dt = s_ds + n * WA_TV;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *vb = vis + (size_t)b * V * D, *mb = mid + (size_t)b * V * H;
    for (int t = tid; t < n * D; t += WA_T) { s_txt[t] = txt[(size_t)b * n * D + t]; s_dt[t] = 0.f; }
    for (int t = tid; t < n * H; t += WA_T) s_do[t] = gout[(size_t)b * n * H + t];
    __syncthreads();
    if (tid < n) {
        s_lse[tid] = lse[(size_t)b * n + tid];
        float dr = 0.f;
        for (int h = 0; h < H; ++h) dr = fmaf(s_do[tid * H + h], out[((size_t)b * n + tid) * H + h], dr);
        s_dr[tid] = dr;
    }
    __syncthreads();
    for (int v0 = 0; v0 < V; v0 += WA_TV) {
        const int tv = min(WA_TV, V - v0);
        for (int t = tid; t < tv * D; t += WA_T) s_vis[(t / D) * DP + t % D] = vb[(size_t)v0 * D + t];
        for (int t = tid; t < tv * H; t += WA_T) s_mid[(t / H) * HP + t % H] = mb[(size_t)v0 * H + t];
        __syncthreads();
        for (int e = tid; e < n * WA_TV; e += WA_T) {
            const int q = e / WA_TV, v = e - q * WA_TV;
            float pv = 0.f, ds = 0.f;
            if (v < tv) {
                float s = 0.f, dp = 0.f;
                const float *a = s_txt + q * D, *c = s_vis + v * DP;
                for (int d = 0; d < D; ++d) s = fmaf(a[d], c[d], s);
                const float *g = s_do + q * H, *m = s_mid + v * HP;
                for (int h = 0; h < H; ++h) dp = fmaf(g[h], m[h], dp);
                pv = __expf(s - s_lse[q]);
                ds = pv * (dp - s_dr[q]);
            }
            s_p[e] = pv;
            s_ds[e] = ds;
        }
        __syncthreads();
        if (gmid && tid < H)  // dmid[v][h] = sum_q P[q][v] dO[q][h]
            for (int v = 0; v < tv; ++v) {
                float a = 0.f;
                for (int q = 0; q < n; ++q) a = fmaf(s_p[q * WA_TV + v], s_do[q * H + tid], a);
                gmid[((size_t)b * V + v0 + v) * H + tid] = a;
            }
        if (gvis)             // dvis[v][d] = sum_q dS[q][v] txt[q][d]
            for (int e = tid; e < tv * D; e += WA_T) {
                const int v = e / D, d = e - v * D;
                float a = 0.f;
                for (int q = 0; q < n; ++q) a = fmaf(s_ds[q * WA_TV + v], s_txt[q * D + d], a);
                gvis[((size_t)b * V + v0) * D + e] = a;
            }
        for (int e = tid; e < n * D; e += WA_T) {  // dtxt[q][d] += sum_v dS[q][v] vis[v][d]
            const int q = e / D, d = e - q * D;
            float a = s_dt[e];
            for (int v = 0; v < tv; ++v) a = fmaf(s_ds[q * WA_TV + v], s_vis[v * DP + d], a);
            s_dt[e] = a;
        }
        __syncthreads();
    }
    if (gtxt)
        for (int t = tid; t < n * D; t += WA_T) gtxt[(size_t)b * n * D + t] = s_dt[t];
}

}  // namespace

cudaError_t launch_word_attention(const float *vis, const float *txt, const float *mid, int B, int V, int n, int D, int H,
                                  float *out, float *lse, cudaStream_t st) {
    if (n > WA_NQ || H > WA_T || n < 1 || D < 1 || H < 1) return cudaErrorInvalidValue;
    const size_t smem = wa_smem<false>(n, D, H);
    cudaError_t e = cudaFuncSetAttribute(word_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    word_attn_fwd_kernel<<<B, WA_T, smem, st>>>(vis, txt, mid, V, n, D, H, out, lse);
    return cudaGetLastError();
}

cudaError_t launch_word_attention_backward(const float *vis, const float *txt, const float *mid, const float *out,
                                           const float *lse, const float *gout, int B, int V, int n, int D, int H, float *gvis,
                                           float *gtxt, float *gmid, cudaStream_t st) {
    if (n > WA_NQ || H > WA_T || n < 1 || D < 1 || H < 1) return cudaErrorInvalidValue;
    const size_t smem = wa_smem<true>(n, D, H);
    cudaError_t e = cudaFuncSetAttribute(word_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    word_attn_bwd_kernel<<<B, WA_T, smem, st>>>(vis, txt, mid, out, lse, gout, V, n, D, H, gvis, gtxt, gmid);
    return cudaGetLastError();
}

}  // namespace vlgae
