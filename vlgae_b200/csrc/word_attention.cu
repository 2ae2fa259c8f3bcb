// word_attention.cu -- per-sample word -> factor attention of DependencyBoxRel._forward.
//
// Reference (/root/reference/src/model/joint.py:668-673):
//     attmap = einsum("bvd,bqd->bqv", vis, txt[:, 1:]).softmax(2);  x = einsum("bqv,bvh->bqh", attmap, vis_mid)
// i.e. one small attention per caption: n <= 64 word queries over V factors (no mask), values vis_mid.  Plain fp32 FMAs
// (0.3 % of the alignment contraction's flops): one CTA per caption, the caption's queries resident in shared memory, the
// factors streamed in tiles of 32 with an online softmax; the [B, n, V] map is never written.  Both small products are
// register-tiled so that an FMA costs a fraction of a shared-memory load: for the scores a thread owns one factor and a
// group of QG queries (operands as 128-bit loads, the query rows broadcast), for the output a thread owns column h for
// every query and reads four probabilities per load.  The backward recomputes the probabilities from the saved row
// log-sum-exp, with the same tiling for its four products.
#include <cuda_runtime.h>
#include <stdint.h>

#include "align_kernels.cuh"

namespace vlgae {
namespace {

constexpr int WA_T = 256;   // threads (= max H)
constexpr int WA_TV = 32;   // factors per tile
constexpr int WA_NQ = 64;   // max queries
// (queries per thread in the score product: template parameter QG = ceil(queries of the CTA / 8 warps), 1..8)

// row strides: D + 4 and H + 4 floats keep rows 16-byte aligned and put the 8 lanes of a 128-bit phase on distinct banks
__host__ __device__ inline int pad4(int x) { return ((x + 3) & ~3) + 4; }

// shared (floats): txt [n][DP] | vis tile [TV][DP] | mid tile [TV][HP] | P [NQ][TV] | m, l, alpha [3 n]
// backward adds:   dO [n][HP] | dS [NQ][TV] | dtxt [n][DP]
template <bool BWD>
__host__ __device__ inline size_t wa_smem(int n, int D, int H) {
    const size_t DP = pad4(D), HP = pad4(H);
    size_t f = (size_t)n * DP + (size_t)WA_TV * DP + (size_t)WA_TV * HP + (size_t)WA_NQ * WA_TV + 3 * (size_t)WA_NQ;
    if (BWD) f += (size_t)n * HP + (size_t)WA_NQ * WA_TV + (size_t)n * DP;
    return f * sizeof(float);
}

// rows [0, nr) of a row-major [*, W] global matrix -> shared rows of stride WP; rows [nr, nr_pad) and the columns from W to
// the next multiple of 4 are zero-filled.  One warp per row, no divisions; 16-byte cp.async (global -> shared without
// registers, all loads of the tile in flight together) when W % 4 == 0 and the source is 16-byte aligned.
__device__ __forceinline__ void load_rows(float *dst, const float *src, int nr, int nr_pad, int W, int WP, int tid) {
    const int warp = tid >> 5, lane = tid & 31, W4 = (W + 3) >> 2;
    const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    for (int r = warp; r < nr_pad; r += WA_T / 32) {
        float *d = dst + r * WP;
        if (r >= nr) {
            for (int c = lane; c < W4; c += 32) reinterpret_cast<float4 *>(d)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (vec) {
            const float *g = src + (size_t)r * W;
            for (int c = lane; c < W4; c += 32) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(d + 4 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g + 4 * c) : "memory");
            }
        } else {
            const float *g = src + (size_t)r * W;
            for (int c = lane; c < 4 * W4; c += 32) d[c] = c < W ? g[c] : 0.f;
        }
    }
}
__device__ __forceinline__ void load_rows_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void load_rows_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// all committed groups but the most recent one have landed
__device__ __forceinline__ void load_rows_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// S[q][v] = <txt[q], vis[v]> for the tile: thread = (factor v = lane, query group = warp); results to s_p[q][v]
template <int WA_QG>
__device__ __forceinline__ void score_tile(const float *s_txt, const float *s_vis, float *s_p, int n, int tv, int D4, int DP, int tid) {
    const int v = tid & 31, q0 = (tid >> 5) * WA_QG;
    if (q0 >= n) return;
    float acc[WA_QG];
#pragma unroll
    for (int k = 0; k < WA_QG; ++k) acc[k] = 0.f;
    const float4 *c = reinterpret_cast<const float4 *>(s_vis + v * DP);
    for (int d = 0; d < D4; ++d) {
        const float4 x = c[d];
#pragma unroll
        for (int k = 0; k < WA_QG; ++k) {
            if (q0 + k < n) {  // warp-uniform
                const float4 a = reinterpret_cast<const float4 *>(s_txt + (q0 + k) * DP)[d];
                acc[k] = fmaf(a.x, x.x, fmaf(a.y, x.y, fmaf(a.z, x.z, fmaf(a.w, x.w, acc[k]))));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < WA_QG; ++k)
        if (q0 + k < n) s_p[(q0 + k) * WA_TV + v] = v < tv ? acc[k] : -INFINITY;
}

// grid (B, query splits): a CTA takes queries [blockIdx.y * nq_cta, ...) of caption blockIdx.x -- several CTAs per SM, so that
// one CTA's tile loads hide behind another's products
template <int WA_QG>
__global__ void __launch_bounds__(WA_T, WA_QG <= 2 ? 3 : 2) word_attn_fwd_kernel(const float *__restrict__ vis, const float *__restrict__ txt,
                                                                const float *__restrict__ mid, int V, int n_all, int nq_cta, int D, int H,
                                                                float *__restrict__ out, float *__restrict__ lse) {
    extern __shared__ __align__(16) float sm[];
    const int qbase = blockIdx.y * nq_cta, n = min(nq_cta, n_all - qbase);
    if (n <= 0) return;
    txt += (size_t)qbase * D; out += (size_t)qbase * H;
    if (lse) lse += qbase;
    const int DP = pad4(D), HP = pad4(H), D4 = (D + 3) >> 2;
    float *s_txt = sm, *s_vis = s_txt + n * DP, *s_mid = s_vis + WA_TV * DP, *s_p = s_mid + WA_TV * HP;
    float *s_m = s_p + WA_NQ * WA_TV, *s_l = s_m + WA_NQ, *s_a = s_l + WA_NQ;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *vb = vis + (size_t)b * V * D, *mb = mid + (size_t)b * V * H;
    // Software pipeline without second buffers: the factor tile is only read by the score product and the value tile only
    // by the output product, so tile t + 1 of each is requested (cp.async) as soon as its consumer of tile t has finished
    // and lands while the other phases run.  Groups are committed in the order vis(t), mid(t), vis(t + 1), ...: "all but
    // the most recent group" is exactly what the next consumer needs.
    load_rows(s_txt, txt + (size_t)b * n_all * D, n, n, D, DP, tid);
    load_rows(s_vis, vb, min(WA_TV, V), min(WA_TV, V), D, DP, tid);
    load_rows_commit();
    load_rows(s_mid, mb, min(WA_TV, V), WA_TV, H, HP, tid);  // rows beyond the tile are zero: their probabilities are zero too
    load_rows_commit();
    for (int q = tid; q < n; q += WA_T) { s_m[q] = -INFINITY; s_l[q] = 0.f; }
    float acc[WA_QG * 8];
#pragma unroll
    for (int q = 0; q < WA_QG * 8; ++q) acc[q] = 0.f;
    for (int v0 = 0; v0 < V; v0 += WA_TV) {
        const int tv = min(WA_TV, V - v0), v1 = v0 + WA_TV, tv1 = min(WA_TV, V - v1);
        load_rows_wait_but_one();  // vis(t)
        __syncthreads();
        score_tile<WA_QG>(s_txt, s_vis, s_p, n, tv, D4, DP, tid);
        __syncthreads();
        if (tv1 > 0) load_rows(s_vis, vb + (size_t)v1 * D, tv1, tv1, D, DP, tid);
        load_rows_commit();
        // online softmax: one warp per query row, one factor per lane
        for (int q = warp; q < n; q += WA_T / 32) {
            const float x = s_p[q * WA_TV + lane];
            float mx = x;
            for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float mo = s_m[q];
            mx = fmaxf(mx, mo);
            const float pv = __expf(x - mx);  // exp(-inf) = 0 beyond the tile
            float sum = pv;
            for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            s_p[q * WA_TV + lane] = pv;
            __syncwarp();  // every lane has read s_m[q] before lane 0 replaces it
            if (lane == 0) {
                const float al = __expf(mo - mx);
                s_l[q] = s_l[q] * al + sum;
                s_m[q] = mx;
                s_a[q] = al;
            }
        }
        load_rows_wait_but_one();  // mid(t)
        __syncthreads();
        if (tid < H) {
            // rescale, then O[q][h] += sum_v P[q][v] mid[v][h]: four factors per probability load
#pragma unroll
            for (int q = 0; q < WA_QG * 8; ++q)
                if (q < n) acc[q] *= s_a[q];
#pragma unroll 1
            for (int v = 0; v < WA_TV; v += 4) {
                const float m0 = s_mid[v * HP + tid], m1 = s_mid[(v + 1) * HP + tid], m2 = s_mid[(v + 2) * HP + tid], m3 = s_mid[(v + 3) * HP + tid];
#pragma unroll
                for (int q = 0; q < WA_QG * 8; ++q) {
                    if (q < n) {
                        const float4 p = *reinterpret_cast<const float4 *>(s_p + q * WA_TV + v);
                        acc[q] = fmaf(p.x, m0, fmaf(p.y, m1, fmaf(p.z, m2, fmaf(p.w, m3, acc[q]))));
                    }
                }
            }
        }
        __syncthreads();
        if (tv1 > 0) load_rows(s_mid, mb + (size_t)v1 * H, tv1, WA_TV, H, HP, tid);
        load_rows_commit();
    }
    load_rows_wait();
    if (tid < H) {
#pragma unroll
        for (int q = 0; q < WA_QG * 8; ++q)
            if (q < n) out[((size_t)b * n_all + q) * H + tid] = acc[q] / s_l[q];
    }
    if (lse && tid < n) lse[(size_t)b * n_all + tid] = s_m[tid] + __logf(s_l[tid]);
}

// backward: P = exp(S - lse);  dmid = P^T dO;  dP = dO mid^T;  dS = P (dP - rowsum(dO * O));  dvis = dS^T txt;  dtxt = dS vis
// grid (B, factor splits): a CTA takes the factors [blockIdx.y * v_cta, ...) of caption blockIdx.x; dvis / dmid rows are its
// own, dtxt is summed over the splits with atomicAdd (zeroed by the launcher when there is more than one split)
template <int WA_QG>
__global__ void __launch_bounds__(WA_T) word_attn_bwd_kernel(const float *__restrict__ vis, const float *__restrict__ txt,
                                                             const float *__restrict__ mid, const float *__restrict__ out,
                                                             const float *__restrict__ lse, const float *__restrict__ gout,
                                                             int V, int v_cta, int n, int D, int H, float *__restrict__ gvis,
                                                             float *__restrict__ gtxt, float *__restrict__ gmid) {
    extern __shared__ __align__(16) float sm[];
    const int DP = pad4(D), HP = pad4(H), D4 = (D + 3) >> 2, H4 = (H + 3) >> 2;
    float *s_txt = sm, *s_vis = s_txt + n * DP, *s_mid = s_vis + WA_TV * DP, *s_p = s_mid + WA_TV * HP;
    float *s_lse = s_p + WA_NQ * WA_TV, *s_dr = s_lse + WA_NQ, *s_unused = s_dr + WA_NQ;
    float *s_do = s_unused + WA_NQ, *s_ds = s_do + n * HP, *s_dt = s_ds + WA_NQ * WA_TV;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *vb = vis + (size_t)b * V * D, *mb = mid + (size_t)b * V * H;
    load_rows(s_txt, txt + (size_t)b * n * D, n, n, D, DP, tid);
    load_rows(s_do, gout + (size_t)b * n * H, n, n, H, HP, tid);
    for (int t = tid; t < n * DP; t += WA_T) s_dt[t] = 0.f;
    load_rows_wait();
    __syncthreads();
    for (int q = warp; q < n; q += WA_T / 32) {  // delta[q] = <dO[q], O[q]>
        float dr = 0.f;
        for (int h = lane; h < H; h += 32) dr = fmaf(s_do[q * HP + h], out[((size_t)b * n + q) * H + h], dr);
        for (int o = 16; o >= 1; o >>= 1) dr += __shfl_xor_sync(0xffffffffu, dr, o);
        if (lane == 0) { s_dr[q] = dr; s_lse[q] = lse[(size_t)b * n + q]; }
    }
    // dO column of this thread for every query (registers), used by the dmid product
    float dcol[WA_QG * 8];
#pragma unroll
    for (int q = 0; q < WA_QG * 8; ++q) dcol[q] = (q < n && tid < H) ? s_do[q * HP + tid] : 0.f;
    const int vlo = blockIdx.y * v_cta, vhi = min(V, vlo + v_cta);
    // same pipeline as the forward kernel: the value tile is free after the dP product and the factor tile after the dtxt
    // product, so the next tiles are requested there and land behind the remaining products of this tile
    if (vlo < vhi) {
        load_rows(s_vis, vb + (size_t)vlo * D, min(WA_TV, vhi - vlo), min(WA_TV, vhi - vlo), D, DP, tid);
        load_rows(s_mid, mb + (size_t)vlo * H, min(WA_TV, vhi - vlo), min(WA_TV, vhi - vlo), H, HP, tid);
    }
    for (int v0 = vlo; v0 < vhi; v0 += WA_TV) {
        const int tv = min(WA_TV, vhi - v0), v1 = v0 + WA_TV, tv1 = min(WA_TV, vhi - v1);
        load_rows_wait();
        __syncthreads();
        score_tile<WA_QG>(s_txt, s_vis, s_p, n, tv, D4, DP, tid);  // S (own entries only: the same thread continues below)
        {
            // dP[q][v] = <dO[q], mid[v]>, same (factor, query group) tiling; then P and dS in place
            const int v = lane, q0 = warp * WA_QG;
            if (q0 < n) {
                float dp[WA_QG];
#pragma unroll
                for (int k = 0; k < WA_QG; ++k) dp[k] = 0.f;
                if (v < tv) {
                    const float4 *c = reinterpret_cast<const float4 *>(s_mid + v * HP);
                    for (int h = 0; h < H4; ++h) {
                        const float4 x = c[h];
#pragma unroll
                        for (int k = 0; k < WA_QG; ++k) {
                            if (q0 + k < n) {
                                const float4 a = reinterpret_cast<const float4 *>(s_do + (q0 + k) * HP)[h];
                                dp[k] = fmaf(a.x, x.x, fmaf(a.y, x.y, fmaf(a.z, x.z, fmaf(a.w, x.w, dp[k]))));
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < WA_QG; ++k) {
                    const int q = q0 + k;
                    if (q < n) {
                        const float pv = v < tv ? __expf(s_p[q * WA_TV + v] - s_lse[q]) : 0.f;
                        s_p[q * WA_TV + v] = pv;
                        s_ds[q * WA_TV + v] = pv * (dp[k] - s_dr[q]);
                    }
                }
            }
        }
        __syncthreads();
        if (tv1 > 0) load_rows(s_mid, mb + (size_t)v1 * H, tv1, tv1, H, HP, tid);
        // dtxt[q][d] += sum_v dS[q][v] vis[v][d]: thread = (4 consecutive d, 2 queries)
        {
            const int nd4 = D4, items = nd4 * ((n + 1) >> 1);
            for (int e = tid; e < items; e += WA_T) {
                const int d4 = e % nd4, q = (e / nd4) * 2;
                const bool two = q + 1 < n;
                float4 a0 = reinterpret_cast<float4 *>(s_dt + q * DP)[d4];
                float4 a1 = two ? reinterpret_cast<float4 *>(s_dt + (q + 1) * DP)[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int v = 0; v < tv; ++v) {
                    const float4 x = reinterpret_cast<const float4 *>(s_vis + v * DP)[d4];
                    const float s0 = s_ds[q * WA_TV + v], s1 = two ? s_ds[(q + 1) * WA_TV + v] : 0.f;
                    a0.x = fmaf(s0, x.x, a0.x); a0.y = fmaf(s0, x.y, a0.y); a0.z = fmaf(s0, x.z, a0.z); a0.w = fmaf(s0, x.w, a0.w);
                    a1.x = fmaf(s1, x.x, a1.x); a1.y = fmaf(s1, x.y, a1.y); a1.z = fmaf(s1, x.z, a1.z); a1.w = fmaf(s1, x.w, a1.w);
                }
                reinterpret_cast<float4 *>(s_dt + q * DP)[d4] = a0;
                if (two) reinterpret_cast<float4 *>(s_dt + (q + 1) * DP)[d4] = a1;
            }
        }
        __syncthreads();
        if (tv1 > 0) load_rows(s_vis, vb + (size_t)v1 * D, tv1, tv1, D, DP, tid);
        if (gmid && tid < H) {  // dmid[v][h] = sum_q P[q][v] dO[q][h]: four factors per probability load
#pragma unroll 1
            for (int v = 0; v < WA_TV; v += 4) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int q = 0; q < WA_QG * 8; ++q) {
                    if (q < n) {
                        const float4 p = *reinterpret_cast<const float4 *>(s_p + q * WA_TV + v);
                        a0 = fmaf(p.x, dcol[q], a0); a1 = fmaf(p.y, dcol[q], a1); a2 = fmaf(p.z, dcol[q], a2); a3 = fmaf(p.w, dcol[q], a3);
                    }
                }
                float *g = gmid + ((size_t)b * V + v0 + v) * H + tid;
                if (v < tv) g[0] = a0;
                if (v + 1 < tv) g[H] = a1;
                if (v + 2 < tv) g[2 * (size_t)H] = a2;
                if (v + 3 < tv) g[3 * (size_t)H] = a3;
            }
        }
        if (gvis) {  // dvis[v][d] = sum_q dS[q][v] txt[q][d]: thread = (4 consecutive d, 4 consecutive v)
            const int nd4 = D4, items = nd4 * (WA_TV / 4);
            for (int e = tid; e < items; e += WA_T) {
                const int d4 = e % nd4, vq = (e / nd4) * 4;
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
                for (int q = 0; q < n; ++q) {
                    const float4 t4 = reinterpret_cast<const float4 *>(s_txt + q * DP)[d4];
                    const float4 s4 = *reinterpret_cast<const float4 *>(s_ds + q * WA_TV + vq);
                    r0.x = fmaf(s4.x, t4.x, r0.x); r0.y = fmaf(s4.x, t4.y, r0.y); r0.z = fmaf(s4.x, t4.z, r0.z); r0.w = fmaf(s4.x, t4.w, r0.w);
                    r1.x = fmaf(s4.y, t4.x, r1.x); r1.y = fmaf(s4.y, t4.y, r1.y); r1.z = fmaf(s4.y, t4.z, r1.z); r1.w = fmaf(s4.y, t4.w, r1.w);
                    r2.x = fmaf(s4.z, t4.x, r2.x); r2.y = fmaf(s4.z, t4.y, r2.y); r2.z = fmaf(s4.z, t4.z, r2.z); r2.w = fmaf(s4.z, t4.w, r2.w);
                    r3.x = fmaf(s4.w, t4.x, r3.x); r3.y = fmaf(s4.w, t4.y, r3.y); r3.z = fmaf(s4.w, t4.z, r3.z); r3.w = fmaf(s4.w, t4.w, r3.w);
                }
                const float4 rr[4] = {r0, r1, r2, r3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (vq + k >= tv) break;
                    float *g = gvis + ((size_t)b * V + v0 + vq + k) * D + d4 * 4;
                    const float vals[4] = {rr[k].x, rr[k].y, rr[k].z, rr[k].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (d4 * 4 + j < D) g[j] = vals[j];
                }
            }
        }
    }
    load_rows_wait();
    __syncthreads();
    if (gtxt)
        for (int t = tid; t < n * D; t += WA_T) {
            const int q = t / D, d = t - q * D;
            if (gridDim.y > 1) atomicAdd(&gtxt[(size_t)b * n * D + t], s_dt[q * DP + d]);
            else gtxt[(size_t)b * n * D + t] = s_dt[q * DP + d];
        }
}

}  // namespace

cudaError_t launch_word_attention(const float *vis, const float *txt, const float *mid, int B, int V, int n, int D, int H,
                                  float *out, float *lse, cudaStream_t st) {
    if (n > WA_NQ || H > WA_T || n < 1 || D < 1 || H < 1) return cudaErrorInvalidValue;
    // query splits: as many as still fit ONE wave of three CTAs per SM (148 SMs; <= 16 queries per CTA keep the registers
    // under the three-CTA bound), at least 8 queries each
    static const int env_slots = [] { const char *v = getenv("VLGAE_WA_SLOTS"); return v && *v ? atoi(v) : 3; }();
    int ns = (env_slots * 148) / (B > 0 ? B : 1);
    if (ns > (n + 7) / 8) ns = (n + 7) / 8;
    if (ns < 1) ns = 1;
    const int nq_cta = (n + ns - 1) / ns;
    ns = (n + nq_cta - 1) / nq_cta;
    const size_t smem = wa_smem<false>(nq_cta, D, H);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<dim3(B, ns), WA_T, smem, st>>>(vis, txt, mid, V, n, nq_cta, D, H, out, lse);
        return cudaGetLastError();
    };
    switch ((nq_cta + 7) / 8) {
        case 1: return go(word_attn_fwd_kernel<1>);
        case 2: return go(word_attn_fwd_kernel<2>);
        case 3: return go(word_attn_fwd_kernel<3>);
        case 4: return go(word_attn_fwd_kernel<4>);
        case 5: return go(word_attn_fwd_kernel<5>);
        case 6: return go(word_attn_fwd_kernel<6>);
        case 7: return go(word_attn_fwd_kernel<7>);
        default: return go(word_attn_fwd_kernel<8>);
    }
}

cudaError_t launch_word_attention_backward(const float *vis, const float *txt, const float *mid, const float *out,
                                           const float *lse, const float *gout, int B, int V, int n, int D, int H, float *gvis,
                                           float *gtxt, float *gmid, cudaStream_t st) {
    if (n > WA_NQ || H > WA_T || n < 1 || D < 1 || H < 1) return cudaErrorInvalidValue;
    // factor splits: enough CTAs for ~2 waves of one CTA per SM, whole tiles each
    int vs = (2 * 148 + B - 1) / B;
    const int tiles = (V + WA_TV - 1) / WA_TV;
    if (vs > tiles) vs = tiles;
    if (vs < 1) vs = 1;
    const int v_cta = ((tiles + vs - 1) / vs) * WA_TV;
    vs = (V + v_cta - 1) / v_cta;
    const size_t smem = wa_smem<true>(n, D, H);
    cudaError_t e = cudaSuccess;
    if (vs > 1 && gtxt && (e = cudaMemsetAsync(gtxt, 0, (size_t)B * n * D * sizeof(float), st)) != cudaSuccess) return e;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e2 != cudaSuccess) return e2;
        kern<<<dim3(B, vs), WA_T, smem, st>>>(vis, txt, mid, out, lse, gout, V, v_cta, n, D, H, gvis, gtxt, gmid);
        return cudaGetLastError();
    };
    switch ((n + 7) / 8) {
        case 1: return go(word_attn_bwd_kernel<1>);
        case 2: return go(word_attn_bwd_kernel<2>);
        case 3: return go(word_attn_bwd_kernel<3>);
        case 4: return go(word_attn_bwd_kernel<4>);
        case 5: return go(word_attn_bwd_kernel<5>);
        case 6: return go(word_attn_bwd_kernel<6>);
        case 7: return go(word_attn_bwd_kernel<7>);
        default: return go(word_attn_bwd_kernel<8>);
    }
}

}  // namespace vlgae
