// dmv_gather.cu -- DMV chart DP, "gather" schedule for the throughput regime (many more sentences than resident CTAs).
//
// Same operator as dmv_frontier.cu (reference /root/reference/src/model/torch_struct/dmv.py:19-66, the autograd
// marginals / argmax of helpers.py:118-154 restated as explicit sweeps), different schedule and chart layout.
//
// Schedule.  One CTA (64 .. 256 threads; one warp for short sentences) owns one sentence.  Widths are swept in order,
// one block barrier per width.  Inside a width every span (i, j = i + w) is evaluated by G = 2^k lanes of one warp, each
// lane streaming a contiguous chunk of the w split points:
//   * log semiring: the reference's own two-pass log-sum-exp (semirings.py:131-132): max over the terms, then
//     sum exp(t - max) in split order, log, + max; the G partial maxima / sums are merged by xor shuffles;
//   * max semiring: running first maximum (strict > in split order; lanes merge on (value, smaller index));
//   * the incomplete items of a span are published with a __syncwarp and the SAME lanes go on to the two complete
//     items of that span (dmv.py:58-62 need IL / IR of the same width for one term only): no barrier in between.
// The reverse sweep (helpers.py:150-154 as an explicit pass) is parent-major: the lanes of span (i, j) push
// g * exp(l + r - parent) into the two operands of each of their terms.  Every accumulator word has one writer per phase
// (phase A': complete parents, dmv.py:58-62 transposed; phase B': incomplete parents, dmv.py:50-56 transposed).
//
// Layout.  Row-major squares of Nb rows x S columns (S even, >= Nb + 1), two triangular item kinds per square, so that
// EVERY operand stream of every step is unit-stride in the split point and its per-lane base advances by S + 1 (odd):
// conflict-free shared-memory accesses and immediate offsets in the unrolled loops (no index arithmetic per term).
//   C  float2  CR[i][r] at (i, r + 1) as (HAS, NO);   CL[j][l] at (j, l) as (NO, HAS)   (swapped: steps 1, 2 pair
//              CR.NO with CL.HAS and CR.HAS with CL.NO, so one packed add forms both terms)
//   I  float2  IR[i][r] at (i, r);  IL[j][l] at (j, l)   (HAS, NO); pre-loaded with attach + dec[GO] (dmv.py:36-37)
//   Ct float   the NOCHILD complete items transposed: CL[h][l].NO at (l, h + 1), CR[h][r].NO at (r, h)
//   X  float   XR(i, j) at (i, j), XL(i, j) at (j, i)   (incomplete items before the arc score; reverse sweep only)
//   gC, gI     gradients, indexed like C and I
//   max pass:  C, I, Ct + first-arg-max bytes BX (like X) and BC (uchar2, like C in natural (HAS, NO) order)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dmv_kernels.cuh"

namespace vlgae {

namespace {

constexpr float NEG_BIG = -1.0e30f;  // below every chart value (the reference's sentinels are -1e12 .. -1e20 sums)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// m + log(s), one rounding of lg2(s) * ln2 + m
__device__ __forceinline__ float lse_fin(float m, float s) { return fmaf(lg2(s), LN2, m); }

template <int NT>
__device__ __forceinline__ void blk_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

struct Geo {
    int Nb, S, S1, E;  // positions, row stride, S + 1, elements per square
};
__host__ __device__ inline int row_stride(int Nb) { return (Nb + 2) & ~1; }  // even, >= Nb + 1

// lanes per span at width w: as many as fit the CTA, never more than the split points allow (>= 2 per lane)
template <int NT>
__device__ __forceinline__ int pick_lg(int n, int w) {
    const int f = NT / n;
    int lg = f >= 1 ? 31 - __clz(f) : 0;
    lg = min(lg, 5);
    const int lw = w >= 2 ? 31 - __clz(w >> 1) : 0;  // floor(log2(w / 2)): chunks of >= 2 split points
    return min(lg, lw);
}

struct Unit {
    int q, il, spw, c, a0, cnt;
};
template <int NT>
__device__ __forceinline__ Unit make_unit(int lane, int lg, int w) {
    Unit u;
    u.spw = 32 >> lg;
    u.q = lane >> (5 - lg);
    u.il = lane & (u.spw - 1);
    u.c = (w + (1 << lg) - 1) >> lg;
    u.a0 = u.q * u.c;
    u.cnt = min(u.c, w - u.a0);  // may be <= 0 for the last lanes of a span
    return u;
}

// ---------------------------------------------------------------------------------------------
// staging: dec, width-0 complete items (STOP decisions, dmv.py:39-40), arc scores attach + dec[GO] (dmv.py:36-37)
// dec index = dir * 4 + val * 2 + decision
// ---------------------------------------------------------------------------------------------
template <int NT, bool WITH_CT>
__device__ __forceinline__ void stage(const DmvArgs &p, int b, const Geo &g, int tid, float *sdec, float2 *C, float2 *I, float *Ct) {
    const int N = p.N, Nb = g.Nb, S = g.S;
    const float *dec = p.dec + (size_t)b * N * 8;
    const float2 *attach = reinterpret_cast<const float2 *>(p.attach + (size_t)b * N * N * 2);
    for (int t = tid; t < Nb * 8; t += NT) sdec[t] = dec[t];
    blk_sync<NT>();
    for (int i = tid; i < Nb; i += NT) {
        const float lh = sdec[i * 8 + 1], ln = sdec[i * 8 + 3], rh = sdec[i * 8 + 5], rn = sdec[i * 8 + 7];
        C[i * S + i] = make_float2(ln, lh);      // CL(i, i) as (NO, HAS)
        C[i * S + i + 1] = make_float2(rh, rn);  // CR(i, i) as (HAS, NO)
        if (WITH_CT) {
            Ct[i * S + i + 1] = ln;  // CLt[i][i]
            Ct[i * S + i] = rn;      // CRt[i][i]
        }
    }
    constexpr int NW = NT / 32;
    const int warp = tid >> 5, lane = tid & 31;
    for (int h = warp; h < Nb; h += NW) {
        const float lg0 = sdec[h * 8 + 0], lg1 = sdec[h * 8 + 2], rg0 = sdec[h * 8 + 4], rg1 = sdec[h * 8 + 6];
        for (int c = lane; c < Nb; c += 32) {
            if (c == h) continue;
            const float2 a = attach[(size_t)h * N + c];
            I[h * S + c] = c < h ? make_float2(__fadd_rn(a.x, lg0), __fadd_rn(a.y, lg1))
                                 : make_float2(__fadd_rn(a.x, rg0), __fadd_rn(a.y, rg1));
        }
    }
}

// Per-word offsets: every arc score into word k is lowered by mu[k], which lowers every tree by sum_k mu[k] -- the
// posteriors are unchanged, log Z is restored at the end -- and keeps the chart values O(10) instead of O(-4 len).
// First guess (attempt 0): best incoming arc + mean STOP costs of the word; attempt 1: the guess corrected by the
// residual zres of the first sweep.  mu[Nb] = sum of the offsets.  Ends with a barrier.
template <int NT>
__device__ __forceinline__ void offsets(const float2 *I, const float *sdec, float *mu, int Nb, int S, int len, int tid,
                                        int attempt, float zres) {
    for (int k = tid; k < Nb; k += NT) {
        float u = 0.f;
        if (attempt == 0) {
            float m = NEG_BIG;
            for (int h = 0; h < Nb; ++h)
                if (h != k) { const float2 v = I[h * S + k]; m = fmaxf(m, fmaxf(v.x, v.y)); }
            if (k >= 1 && m > -1e6f)
                u = m + 0.5f * (sdec[k * 8 + 1] + sdec[k * 8 + 3]) + 0.5f * (sdec[k * 8 + 5] + sdec[k * 8 + 7]);
            if (!(fabsf(u) < 1e6f)) u = 0.f;
        } else if (k >= 1) {
            u = mu[k] + zres / (float)len;
        }
        mu[k] = u;
    }
    blk_sync<NT>();
    if (tid < 32) {
        float part = 0.f;
        for (int k = 1 + tid; k < Nb; k += 32) part += mu[k];
        for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (tid == 0) mu[Nb] = part;
    }
    blk_sync<NT>();
}

// ---------------------------------------------------------------------------------------------
// log semiring in the log domain (two-pass log-sum-exp): the fallback of lin_pass
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ void log_pass(const DmvArgs &p, int b, int len, unsigned char *smem, int tid) {
    Geo g;
    g.Nb = len + 1; g.S = row_stride(g.Nb); g.S1 = g.S + 1; g.E = g.Nb * g.S;
    const int Nb = g.Nb, S = g.S, S1 = g.S1, E = g.E, N = p.N;
    float2 *C = reinterpret_cast<float2 *>(smem);
    float2 *I = C + E, *gC = I + E, *gI = gC + E;
    float *Ct = reinterpret_cast<float *>(gI + E), *X = Ct + E, *sdec = X + E;
    const int warp = NT == 32 ? 0 : (tid >> 5), lane = tid & 31;
    constexpr int NW = NT / 32;
    const bool want_grad = (p.gdec != nullptr) || (p.gattach != nullptr);

    float *mu = sdec + Nb * 8;  // Nb per-word offsets + their sum
    float zres = 0.f;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
    stage<NT, true>(p, b, g, tid, sdec, C, I, Ct);
    if (want_grad && attempt == 0) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *g4 = reinterpret_cast<float4 *>(gC);  // gC and gI are contiguous: E float4
        for (int t = tid; t < E; t += NT) g4[t] = z;
    }
    blk_sync<NT>();
    // Per-word offsets (see shift_arcs): every arc score into word k is lowered by mu[k], which lowers every tree by
    // sum_k mu[k] -- posteriors unchanged, log Z restored at the end -- and keeps the chart values O(10) instead of
    // O(-4 len), where one fp32 ulp is 1e-6 instead of 1.5e-5.  First guess: best incoming arc + mean STOP costs of the
    // word; if the top of the chart still ends up far from 0 the sweep is repeated once with the guess corrected.
    for (int k = tid; k < Nb; k += NT) {
        float u = 0.f;
        if (attempt == 0) {
            float m = NEG_BIG;
            for (int h = 0; h < Nb; ++h)
                if (h != k) { const float2 v = I[h * S + k]; m = fmaxf(m, fmaxf(v.x, v.y)); }
            if (k >= 1 && m > -1e6f)
                u = m + 0.5f * (sdec[k * 8 + 1] + sdec[k * 8 + 3]) + 0.5f * (sdec[k * 8 + 5] + sdec[k * 8 + 7]);
            if (!(fabsf(u) < 1e6f)) u = 0.f;
        } else if (k >= 1) {
            u = mu[k] + zres / (float)len;
        }
        mu[k] = u;
    }
    blk_sync<NT>();
    if (tid == 0) {
        float t = 0.f;
        for (int k = 1; k < Nb; ++k) t += mu[k];
        mu[Nb] = t;
    }
    for (int h = warp; h < Nb; h += NW)
        for (int c = lane; c < Nb; c += 32) {
            if (c == h) continue;
            const float2 v = I[h * S + c];
            const float m = mu[c];
            I[h * S + c] = make_float2(__fadd_rn(v.x, -m), __fadd_rn(v.y, -m));
        }
    blk_sync<NT>();

    // ---------------- inside ----------------
#pragma unroll 1
    for (int w = 1; w <= len; ++w) {
        const int n = Nb - w;
        const int lg = pick_lg<NT>(n, w);
        const Unit u = make_unit<NT>(lane, lg, w);
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            const bool valid = i < n;
            const int ic = valid ? i : n - 1, j = ic + w;
            // steps 1, 2 (dmv.py:50-56): XL = (+)_r CR[i,r].NO + CL[j,r+1].HAS, XR = (+)_r CR[i,r].HAS + CL[j,r+1].NO
            const float2 *pL = C + ic * S1 + 1 + u.a0;     // CR(i, i + a)       (HAS, NO)
            const float2 *pR = C + j * S + ic + 1 + u.a0;  // CL(j, i + 1 + a)   (NO, HAS)
            float2 m = make_float2(NEG_BIG, NEG_BIG);      // (XR, XL)
#pragma unroll 4
            for (int k = 0; k < u.cnt; ++k) {
                const float2 t = __fadd2_rn(pL[k], pR[k]);
                m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y);
            }
            for (int o = 16; o >= u.spw; o >>= 1) {
                m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, o));
                m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, o));
            }
            const float2 nm = make_float2(-m.x, -m.y), l2 = make_float2(LOG2E, LOG2E);
            float2 s = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int k = 0; k < u.cnt; ++k) {
                const float2 t = __fadd2_rn(pL[k], pR[k]);
                const float2 d = __fmul2_rn(__fadd2_rn(t, nm), l2);
                s = __fadd2_rn(s, make_float2(ex2(d.x), ex2(d.y)));
            }
            for (int o = 16; o >= u.spw; o >>= 1) {
                s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
                s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            }
            const float xr = lse_fin(m.x, s.x), xl = lse_fin(m.y, s.y);
            if (valid && u.q == 0) {
                float2 *il = I + j * S + ic, *ir = I + ic * S + j;
                const float2 al = *il, ar = *ir;
                *il = make_float2(__fadd_rn(xl, al.x), __fadd_rn(xl, al.y));
                *ir = make_float2(__fadd_rn(xr, ar.x), __fadd_rn(xr, ar.y));
                X[j * S + ic] = xl;
                X[ic * S + j] = xr;
            }
            __syncwarp();
            // step 3 (dmv.py:58-59): CL[j,i][v] = (+)_r CL[r,i].NO + IL[j,r][v],  r = i + a
            // step 4 (dmv.py:61-62): CR[i,j][v] = (+)_r IR[i,r][v] + CR[r,j].NO,  r = i + 1 + a
            const float *tL = Ct + ic * S1 + 1 + u.a0;     // CLt[i][i + a]
            const float2 *iL = I + j * S + ic + u.a0;      // IL(j, i + a)
            const float2 *iR = I + ic * S1 + 1 + u.a0;     // IR(i, i + 1 + a)
            const float *tR = Ct + j * S + ic + 1 + u.a0;  // CRt[j][i + 1 + a]
            float2 ml = make_float2(NEG_BIG, NEG_BIG), mr = ml;
#pragma unroll 4
            for (int k = 0; k < u.cnt; ++k) {
                const float l3 = tL[k], r4 = tR[k];
                const float2 i3 = iL[k], i4 = iR[k];
                ml.x = fmaxf(ml.x, __fadd_rn(l3, i3.x)); ml.y = fmaxf(ml.y, __fadd_rn(l3, i3.y));
                mr.x = fmaxf(mr.x, __fadd_rn(i4.x, r4)); mr.y = fmaxf(mr.y, __fadd_rn(i4.y, r4));
            }
            for (int o = 16; o >= u.spw; o >>= 1) {
                ml.x = fmaxf(ml.x, __shfl_xor_sync(0xffffffffu, ml.x, o));
                ml.y = fmaxf(ml.y, __shfl_xor_sync(0xffffffffu, ml.y, o));
                mr.x = fmaxf(mr.x, __shfl_xor_sync(0xffffffffu, mr.x, o));
                mr.y = fmaxf(mr.y, __shfl_xor_sync(0xffffffffu, mr.y, o));
            }
            const float2 nml = make_float2(-ml.x, -ml.y), nmr = make_float2(-mr.x, -mr.y);
            float2 sl = make_float2(0.f, 0.f), sr = sl;
#pragma unroll 4
            for (int k = 0; k < u.cnt; ++k) {
                const float l3 = tL[k], r4 = tR[k];
                const float2 i3 = iL[k], i4 = iR[k];
                const float2 tl = make_float2(__fadd_rn(l3, i3.x), __fadd_rn(l3, i3.y));
                const float2 tr = make_float2(__fadd_rn(i4.x, r4), __fadd_rn(i4.y, r4));
                const float2 dl = __fmul2_rn(__fadd2_rn(tl, nml), l2), dr = __fmul2_rn(__fadd2_rn(tr, nmr), l2);
                sl = __fadd2_rn(sl, make_float2(ex2(dl.x), ex2(dl.y)));
                sr = __fadd2_rn(sr, make_float2(ex2(dr.x), ex2(dr.y)));
            }
            for (int o = 16; o >= u.spw; o >>= 1) {
                sl.x += __shfl_xor_sync(0xffffffffu, sl.x, o);
                sl.y += __shfl_xor_sync(0xffffffffu, sl.y, o);
                sr.x += __shfl_xor_sync(0xffffffffu, sr.x, o);
                sr.y += __shfl_xor_sync(0xffffffffu, sr.y, o);
            }
            if (valid && u.q == 0) {
                const float2 cl = make_float2(lse_fin(ml.x, sl.x), lse_fin(ml.y, sl.y));
                float2 cr = make_float2(lse_fin(mr.x, sr.x), lse_fin(mr.y, sr.y));
                if (ic == 0 && w != len) cr = make_float2(p.mask_zero, p.mask_zero);  // single-root mask, dmv.py:63
                C[j * S + ic] = make_float2(cl.y, cl.x);
                C[ic * S + j + 1] = cr;
                Ct[ic * S + j + 1] = cl.y;
                Ct[j * S + ic] = cr.y;
            }
        }
        blk_sync<NT>();
    }
    zres = C[len + 1].y;  // CR(0, len).NO (dmv.py:65) minus the offsets
    if (fabsf(zres) <= 48.f || len == 0) break;
    blk_sync<NT>();
    }
    if (tid == 0) p.Z[b] = zres + mu[Nb];
    if (!want_grad) { blk_sync<NT>(); return; }

    // ---------------- outside (explicit reverse sweep) ----------------
    if (tid == 0) gC[len + 1].y = p.gZ ? p.gZ[b] : 1.f;
    blk_sync<NT>();
    const float2 l2 = make_float2(LOG2E, LOG2E);
#pragma unroll 1
    for (int w = len; w >= 1; --w) {
        const int n = Nb - w;
        const int lg = pick_lg<NT>(n, w);
        const Unit u = make_unit<NT>(lane, lg, w);
        // phase A'(w): complete parents of width w (steps 3, 4 transposed)
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            if (i >= n || u.cnt <= 0) continue;
            const int j = i + w;
            const float2 plv = C[j * S + i], plg = gC[j * S + i];          // CL(j, i): (NO, HAS)
            const float2 prv = C[i * S + j + 1];                           // CR(i, j): (HAS, NO)
            float2 prg = gC[i * S + j + 1];
            const float2 ncl = make_float2(-plv.y, -plv.x), gcl = make_float2(plg.y, plg.x);  // (HAS, NO)
            float2 ncr = make_float2(-prv.x, -prv.y);
            // the mask overwrote CR[0][w] (dmv.py:63): no gradient passes (and t - mask_zero must not overflow the exp)
            if (i == 0 && w != len) { prg = make_float2(0.f, 0.f); ncr = make_float2(NEG_BIG, NEG_BIG); }
            const float *tL = Ct + i * S1 + 1 + u.a0;
            float2 *iL = I + j * S + i + u.a0, *giL = gI + j * S + i + u.a0;
            float2 *iR = I + i * S1 + 1 + u.a0, *giR = gI + i * S1 + 1 + u.a0;
            const float *tR = Ct + j * S + i + 1 + u.a0;
            float *gl = &gC[(i + u.a0) * S + i].x;          // gCL(i + a, i).NO
            float *gr = &gC[(i + 1 + u.a0) * S + j + 1].y;  // gCR(i + 1 + a, j).NO
            const int gs = 2 * S;
#pragma unroll 2
            for (int k = 0; k < u.cnt; ++k) {
                const float l3 = tL[k], r4 = tR[k];
                const float2 i3 = iL[k], i4 = iR[k];
                const float2 o3 = giL[k], o4 = giR[k];
                const float ol = gl[k * gs], orr = gr[k * gs];
                const float2 tl = make_float2(__fadd_rn(l3, i3.x), __fadd_rn(l3, i3.y));
                const float2 tr = make_float2(__fadd_rn(i4.x, r4), __fadd_rn(i4.y, r4));
                const float2 dl = __fmul2_rn(__fadd2_rn(tl, ncl), l2), dr = __fmul2_rn(__fadd2_rn(tr, ncr), l2);
                const float2 pl = __fmul2_rn(gcl, make_float2(ex2(dl.x), ex2(dl.y)));
                const float2 pr = __fmul2_rn(prg, make_float2(ex2(dr.x), ex2(dr.y)));
                giL[k] = __fadd2_rn(o3, pl);
                giR[k] = __fadd2_rn(o4, pr);
                gl[k * gs] = ol + (pl.x + pl.y);
                gr[k * gs] = orr + (pr.x + pr.y);
            }
        }
        blk_sync<NT>();
        // phase B'(w): incomplete parents of width w (steps 1, 2 transposed)
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            if (i >= n || u.cnt <= 0) continue;
            const int j = i + w;
            const float2 gil = gI[j * S + i], gir = gI[i * S + j];
            const float2 gx = make_float2(gir.x + gir.y, gil.x + gil.y);  // (gXR, gXL)
            const float2 nx = make_float2(-X[i * S + j], -X[j * S + i]);  // -(XR, XL)
            const float2 *pL = C + i * S1 + 1 + u.a0, *pR = C + j * S + i + 1 + u.a0;
            float2 *qL = gC + i * S1 + 1 + u.a0, *qR = gC + j * S + i + 1 + u.a0;
#pragma unroll 2
            for (int k = 0; k < u.cnt; ++k) {
                const float2 a = pL[k], c2 = pR[k], oa = qL[k], oc = qR[k];
                const float2 d = __fmul2_rn(__fadd2_rn(__fadd2_rn(a, c2), nx), l2);
                const float2 pp = __fmul2_rn(gx, make_float2(ex2(d.x), ex2(d.y)));
                qL[k] = __fadd2_rn(oa, pp);  // CR(i, r): .HAS from step 2, .NO from step 1
                qR[k] = __fadd2_rn(oc, pp);  // CL(j, r + 1) stored (NO, HAS): .NO from step 2, .HAS from step 1
            }
        }
        blk_sync<NT>();
    }

    // ---------------- outputs ----------------
    if (p.gattach) {
        float2 *ga = reinterpret_cast<float2 *>(p.gattach + (size_t)b * N * N * 2);
        for (int h = warp; h < N; h += NW)
            for (int c = lane; c < N; c += 32) {
                float2 v = make_float2(0.f, 0.f);
                if (h < Nb && c < Nb && h != c) v = gI[h * S + c];
                ga[(size_t)h * N + c] = v;
            }
    }
    if (p.gdec) {
        float *gd = p.gdec + (size_t)b * N * 8;
        for (int t = tid; t < N * 2; t += NT) {
            const int i = t >> 1, dir = t & 1;
            float2 go = make_float2(0.f, 0.f), stop = make_float2(0.f, 0.f);
            if (i < Nb) {
                if (dir == 0) {
                    for (int c = 0; c < i; ++c) { const float2 v = gI[i * S + c]; go.x += v.x; go.y += v.y; }
                    const float2 sv = gC[i * S + i];  // (NO, HAS)
                    stop = make_float2(sv.y, sv.x);
                } else {
                    for (int c = i + 1; c < Nb; ++c) { const float2 v = gI[i * S + c]; go.x += v.x; go.y += v.y; }
                    stop = gC[i * S + i + 1];
                }
            }
            *reinterpret_cast<float4 *>(gd + i * 8 + dir * 4) = make_float4(go.x, stop.x, go.y, stop.y);  // [dir][val][decision]
        }
    }
    blk_sync<NT>();
}

// ---------------------------------------------------------------------------------------------
// log semiring in the LINEAR domain (default).  With the per-word offsets every chart value exp(v - offsets) is
// O(e^+-20), well inside the fp32 range, so the log-semiring recurrences (dmv.py:50-62 with logsumexp / +) can be run as
// plain sums of products: one FMA per split-point term, no exp, no max pass, no log.  The reverse sweep propagates
// linear adjoints: for alpha_P += alpha_L alpha_R,  beta_L += beta_P alpha_R and beta_R += beta_P alpha_L; seeded with
// gZ / Z', alpha_item * beta_item is d(gZ log Z) / d(log potential) -- the arc marginals and decision counts the
// reference obtains by autograd through the chart (helpers.py:150-154).
// The fused width step: for span (i, j) the lanes stream a in [0, w-1) once and accumulate, in the same loop, the terms of
// X(a), of CL with split a + 1 and of CR with split a (all operands narrower than w); the one remaining X term and the
// two same-span terms (CL split 0 needs IL(j,i), CR split w-1 needs IR(i,j)) are added after the single xor-shuffle
// reduction -- no __syncwarp, one reduction and one barrier per width.
// Returns false when the sweep left the fp32 range or fails its self-check (every word's arc marginals must sum to gZ):
// the caller then runs the log-domain sweep (log_pass) for this sentence.
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ bool blk_all(bool ok) {
    if (NT == 32) return __all_sync(0xffffffffu, ok);
    return __syncthreads_and(ok);
}

template <int NT>
__device__ bool lin_pass(const DmvArgs &p, int b, int len, unsigned char *smem, int tid) {
    Geo g;
    g.Nb = len + 1; g.S = row_stride(g.Nb); g.S1 = g.S + 1; g.E = g.Nb * g.S;
    const int Nb = g.Nb, S = g.S, S1 = g.S1, E = g.E, N = p.N;
    float2 *C = reinterpret_cast<float2 *>(smem);
    float2 *I = C + E, *gC = I + E, *gI = gC + E;
    float *Ct = reinterpret_cast<float *>(gI + E), *X = Ct + E, *sdec = X + E;
    const int warp = NT == 32 ? 0 : (tid >> 5), lane = tid & 31;
    constexpr int NW = NT / 32;
    const bool want_grad = (p.gdec != nullptr) || (p.gattach != nullptr);
    float *mu = sdec + Nb * 8;  // Nb per-word offsets + their sum
    float zres = 0.f, ztop = 1.f;
    bool ok = true;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        stage<NT, true>(p, b, g, tid, sdec, C, I, Ct);
        if (want_grad && attempt == 0) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 *g4 = reinterpret_cast<float4 *>(gC);  // gC and gI are contiguous: E float4
            for (int t = tid; t < E; t += NT) g4[t] = z;
        }
        blk_sync<NT>();
        offsets<NT>(I, sdec, mu, Nb, S, len, tid, attempt, zres);
        // to the linear domain: STOP factors as they are, arc factors lowered by the offset of their child
        for (int i = tid; i < Nb; i += NT) {
            const float2 cl = C[i * S + i], cr = C[i * S + i + 1];
            const float2 el = make_float2(ex2(cl.x * LOG2E), ex2(cl.y * LOG2E)), er = make_float2(ex2(cr.x * LOG2E), ex2(cr.y * LOG2E));
            C[i * S + i] = el; C[i * S + i + 1] = er;
            Ct[i * S + i + 1] = el.x; Ct[i * S + i] = er.y;
        }
        for (int h = warp; h < Nb; h += NW)
            for (int c = lane; c < Nb; c += 32) {
                if (c == h) continue;
                const float2 v = I[h * S + c];
                const float m = mu[c];
                I[h * S + c] = make_float2(ex2(__fadd_rn(v.x, -m) * LOG2E), ex2(__fadd_rn(v.y, -m) * LOG2E));
            }
        blk_sync<NT>();

        // ---------------- inside ----------------
#pragma unroll 1
        for (int w = 1; w <= len; ++w) {
            const int n = Nb - w;
            const int lg = pick_lg<NT>(n, w);
            const Unit u = make_unit<NT>(lane, lg, w - 1);  // the fused loop covers a in [0, w - 1)
#pragma unroll 1
            for (int base = warp * u.spw; base < n; base += NW * u.spw) {
                const int i = base + u.il;
                const bool valid = i < n;
                const int ic = valid ? i : n - 1, j = ic + w;
                const float2 *pL = C + ic * S1 + 1 + u.a0;     // CR(i, i + a)          (HAS, NO)
                const float2 *pR = C + j * S + ic + 1 + u.a0;  // CL(j, i + 1 + a)      (NO, HAS)
                const float *tL = Ct + ic * S1 + 2 + u.a0;     // CLt[i][i + a + 1]     split a + 1 of step 3
                const float2 *iL = I + j * S + ic + 1 + u.a0;  // IL(j, i + a + 1)
                const float2 *iR = I + ic * S1 + 1 + u.a0;     // IR(i, i + 1 + a)      split a of step 4
                const float *tR = Ct + j * S + ic + 1 + u.a0;  // CRt[j][i + 1 + a]
                float2 x = make_float2(0.f, 0.f), cl = x, cr = x;  // x = (XR, XL)
#pragma unroll 4
                for (int k = 0; k < u.cnt; ++k) {
                    x = __ffma2_rn(pL[k], pR[k], x);
                    const float l3 = tL[k], r4 = tR[k];
                    const float2 i3 = iL[k], i4 = iR[k];
                    cl.x = fmaf(l3, i3.x, cl.x); cl.y = fmaf(l3, i3.y, cl.y);
                    cr.x = fmaf(i4.x, r4, cr.x); cr.y = fmaf(i4.y, r4, cr.y);
                }
                if (u.q == 0) x = __ffma2_rn(C[ic * S + ic + w], C[j * S + j], x);  // X term a = w - 1: CR(i, j-1) CL(j, j)
                for (int o = 16; o >= u.spw; o >>= 1) {
                    x.x += __shfl_xor_sync(0xffffffffu, x.x, o); x.y += __shfl_xor_sync(0xffffffffu, x.y, o);
                    cl.x += __shfl_xor_sync(0xffffffffu, cl.x, o); cl.y += __shfl_xor_sync(0xffffffffu, cl.y, o);
                    cr.x += __shfl_xor_sync(0xffffffffu, cr.x, o); cr.y += __shfl_xor_sync(0xffffffffu, cr.y, o);
                }
                if (valid && u.q == 0) {
                    float2 *pil = I + j * S + ic, *pir = I + ic * S + j;
                    const float2 al = *pil, ar = *pir;
                    const float2 il = make_float2(x.y * al.x, x.y * al.y), ir = make_float2(x.x * ar.x, x.x * ar.y);
                    const float l0 = Ct[ic * S1 + 1], r0 = Ct[j * S + j];  // CL(i,i).NO, CR(j,j).NO
                    cl.x = fmaf(l0, il.x, cl.x); cl.y = fmaf(l0, il.y, cl.y);   // step 3, split 0
                    cr.x = fmaf(ir.x, r0, cr.x); cr.y = fmaf(ir.y, r0, cr.y);   // step 4, split w - 1
                    if (ic == 0 && w != len) cr = make_float2(0.f, 0.f);        // single-root mask, dmv.py:63 (exp(-1e12))
                    *pil = il; *pir = ir;
                    X[j * S + ic] = x.y; X[ic * S + j] = x.x;
                    C[j * S + ic] = make_float2(cl.y, cl.x);
                    C[ic * S + j + 1] = cr;
                    Ct[ic * S + j + 1] = cl.y;
                    Ct[j * S + ic] = cr.y;
                }
            }
            blk_sync<NT>();
        }
        ztop = C[len + 1].y;  // CR(0, len).NO (dmv.py:65) in the linear domain
        ok = ztop > 1e-30f && ztop < 1e30f;  // also false for NaN
        if (!ok) break;
        zres = lg2(ztop) * LN2;
        if (fabsf(zres) <= 16.f || len == 0) break;
        blk_sync<NT>();
    }
    if (!ok) { blk_sync<NT>(); return false; }
    if (!want_grad) {
        if (tid == 0) p.Z[b] = zres + mu[Nb];
        blk_sync<NT>();
        return true;
    }

    // ---------------- outside (linear adjoints) ----------------
    const float gz = p.gZ ? p.gZ[b] : 1.f;
    if (tid == 0) gC[len + 1].y = __fdividef(gz, ztop);
    blk_sync<NT>();
#pragma unroll 1
    for (int w = len; w >= 1; --w) {
        const int n = Nb - w;
        const int lg = pick_lg<NT>(n, w);
        const Unit u = make_unit<NT>(lane, lg, w);
        // phase A'(w): complete parents of width w (steps 3, 4 transposed)
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            if (i >= n || u.cnt <= 0) continue;
            const int j = i + w;
            const float2 plg = gC[j * S + i];  // beta CL(j, i): (NO, HAS)
            float2 prg = gC[i * S + j + 1];    // beta CR(i, j): (HAS, NO)
            if (i == 0 && w != len) prg = make_float2(0.f, 0.f);  // the mask overwrote CR[0][w]: nothing passes
            const float2 bcl = make_float2(plg.y, plg.x);
            const float *tL = Ct + i * S1 + 1 + u.a0;
            const float2 *iL = I + j * S + i + u.a0, *iR = I + i * S1 + 1 + u.a0;
            float2 *giL = gI + j * S + i + u.a0, *giR = gI + i * S1 + 1 + u.a0;
            const float *tR = Ct + j * S + i + 1 + u.a0;
            float *gl = &gC[(i + u.a0) * S + i].x;          // beta CL(i + a, i).NO
            float *gr = &gC[(i + 1 + u.a0) * S + j + 1].y;  // beta CR(i + 1 + a, j).NO
            const int gs = 2 * S;
#pragma unroll 2
            for (int k = 0; k < u.cnt; ++k) {
                const float l3 = tL[k], r4 = tR[k];
                const float2 i3 = iL[k], i4 = iR[k];
                const float2 o3 = giL[k], o4 = giR[k];
                const float ol = gl[k * gs], orr = gr[k * gs];
                giL[k] = make_float2(fmaf(bcl.x, l3, o3.x), fmaf(bcl.y, l3, o3.y));
                giR[k] = make_float2(fmaf(prg.x, r4, o4.x), fmaf(prg.y, r4, o4.y));
                gl[k * gs] = fmaf(bcl.x, i3.x, fmaf(bcl.y, i3.y, ol));
                gr[k * gs] = fmaf(prg.x, i4.x, fmaf(prg.y, i4.y, orr));
            }
        }
        blk_sync<NT>();
        // phase B'(w): incomplete parents of width w (steps 1, 2 transposed); beta X = sum_v beta I(v) E(v), E = I / X
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            if (i >= n || u.cnt <= 0) continue;
            const int j = i + w;
            const float2 gil = gI[j * S + i], gir = gI[i * S + j], ail = I[j * S + i], air = I[i * S + j];
            const float xl = X[j * S + i], xr = X[i * S + j];
            const float2 bx = make_float2(xr > 0.f ? __fdividef(fmaf(gir.x, air.x, gir.y * air.y), xr) : 0.f,
                                          xl > 0.f ? __fdividef(fmaf(gil.x, ail.x, gil.y * ail.y), xl) : 0.f);  // (XR, XL)
            const float2 *pL = C + i * S1 + 1 + u.a0, *pR = C + j * S + i + 1 + u.a0;
            float2 *qL = gC + i * S1 + 1 + u.a0, *qR = gC + j * S + i + 1 + u.a0;
#pragma unroll 2
            for (int k = 0; k < u.cnt; ++k) {
                const float2 a = pL[k], c2 = pR[k], oa = qL[k], oc = qR[k];
                qL[k] = __ffma2_rn(bx, c2, oa);  // CR(i, r): .HAS from step 2, .NO from step 1
                qR[k] = __ffma2_rn(bx, a, oc);   // CL(j, r + 1) stored (NO, HAS): .NO from step 2, .HAS from step 1
            }
        }
        blk_sync<NT>();
    }

    // ---------------- outputs: alpha * beta, self-check, copy out ----------------
    for (int h = warp; h < Nb; h += NW)
        for (int c = lane; c < Nb; c += 32) {
            if (c == h) continue;
            const float2 a = I[h * S + c], bt = gI[h * S + c];
            gI[h * S + c] = make_float2(a.x == 0.f ? 0.f : a.x * bt.x, a.y == 0.f ? 0.f : a.y * bt.y);
        }
    blk_sync<NT>();
    bool good = true;
    for (int c = 1 + tid; c < Nb; c += NT) {  // every word has exactly one head: its arc marginals sum to gZ
        float t = 0.f;
        for (int h = 0; h < Nb; ++h)
            if (h != c) { const float2 v = gI[h * S + c]; t += v.x + v.y; }
        if (!(fabsf(t - gz) <= 1e-3f * fabsf(gz) + 1e-30f)) good = false;
    }
    if (!blk_all<NT>(good)) return false;
    if (tid == 0) p.Z[b] = zres + mu[Nb];
    if (p.gattach) {
        float2 *ga = reinterpret_cast<float2 *>(p.gattach + (size_t)b * N * N * 2);
        for (int h = warp; h < N; h += NW)
            for (int c = lane; c < N; c += 32) {
                float2 v = make_float2(0.f, 0.f);
                if (h < Nb && c < Nb && h != c) v = gI[h * S + c];
                ga[(size_t)h * N + c] = v;
            }
    }
    if (p.gdec) {
        float *gd = p.gdec + (size_t)b * N * 8;
        for (int t = tid; t < N * 2; t += NT) {
            const int i = t >> 1, dir = t & 1;
            float2 go = make_float2(0.f, 0.f), stop = make_float2(0.f, 0.f);
            if (i < Nb) {
                if (dir == 0) {
                    for (int c = 0; c < i; ++c) { const float2 v = gI[i * S + c]; go.x += v.x; go.y += v.y; }
                    const float2 a = C[i * S + i], bt = gC[i * S + i];  // (NO, HAS)
                    stop = make_float2(a.y * bt.y, a.x * bt.x);
                } else {
                    for (int c = i + 1; c < Nb; ++c) { const float2 v = gI[i * S + c]; go.x += v.x; go.y += v.y; }
                    const float2 a = C[i * S + i + 1], bt = gC[i * S + i + 1];
                    stop = make_float2(a.x * bt.x, a.y * bt.y);
                }
            }
            *reinterpret_cast<float4 *>(gd + i * 8 + dir * 4) = make_float4(go.x, stop.x, go.y, stop.y);  // [dir][val][decision]
        }
    }
    blk_sync<NT>();
    return true;
}

// ---------------------------------------------------------------------------------------------
// max semiring: Viterbi chart with first-max back-pointers + breadth-first back-trace
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void amax(float &v, int &a, float t, int idx) {
    if (t > v) { v = t; a = idx; }
}
__device__ __forceinline__ void amerge(float &v, int &a, int o) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oa = __shfl_xor_sync(0xffffffffu, a, o);
    if (ov > v || (ov == v && oa < a)) { v = ov; a = oa; }
}
// items of the back-trace: kind (0 CR, 1 CL, 2 IR, 3 IL) | v << 2 | lo << 3 | hi << 12
__device__ __forceinline__ int mk_item(int kind, int v, int lo, int hi) { return kind | (v << 2) | (lo << 3) | (hi << 12); }

template <int NT>
__device__ void max_pass(const DmvArgs &p, int b, int len, unsigned char *smem, int tid) {
    Geo g;
    g.Nb = len + 1; g.S = row_stride(g.Nb); g.S1 = g.S + 1; g.E = g.Nb * g.S;
    const int Nb = g.Nb, S = g.S, S1 = g.S1, E = g.E, N = p.N;
    float2 *C = reinterpret_cast<float2 *>(smem);
    float2 *I = C + E;
    float *Ct = reinterpret_cast<float *>(I + E), *sdec = Ct + E;
    int *queue = reinterpret_cast<int *>(sdec + Nb * 8);  // 2 x (2 Nb + 2) ints
    uchar2 *BC = reinterpret_cast<uchar2 *>(queue + 2 * (2 * Nb + 2));
    uint8_t *BX = reinterpret_cast<uint8_t *>(BC + E);
    const int warp = NT == 32 ? 0 : (tid >> 5), lane = tid & 31;
    constexpr int NW = NT / 32;

    stage<NT, true>(p, b, g, tid, sdec, C, I, Ct);
    // outputs that the back-trace only dots with ones are zero-filled up front
    if (p.arcs) {
        float2 *z = reinterpret_cast<float2 *>(p.arcs + (size_t)b * N * N * 2);
        for (int t = tid; t < N * N; t += NT) z[t] = make_float2(0.f, 0.f);
    }
    if (p.vgdec) for (int t = tid; t < N * 8; t += NT) p.vgdec[(size_t)b * N * 8 + t] = 0.f;
    if (p.heads) for (int t = tid; t < N; t += NT) p.heads[(size_t)b * N + t] = 0;
    blk_sync<NT>();

    // Fused width step (as in lin_pass): the lanes of span (i, j) stream a in [0, w-1) once and fold, in the same loop,
    // the terms of X(a), of CL with split a + 1 and of CR with split a; the remaining X term (a = w-1) and the two
    // same-span terms are folded after the single shuffle merge.  First-maximum rule (torch.max returns the smallest
    // maximal index, semirings.py:200-202): strict > inside a lane's increasing chunk, (value, smaller index) across
    // lanes, CL's split 0 wins ties (>=), CR's split w-1 and X's a = w-1 lose them (>).
#pragma unroll 1
    for (int w = 1; w <= len; ++w) {
        const int n = Nb - w;
        const int lg = pick_lg<NT>(n, w);
        const Unit u = make_unit<NT>(lane, lg, w - 1);
#pragma unroll 1
        for (int base = warp * u.spw; base < n; base += NW * u.spw) {
            const int i = base + u.il;
            const bool valid = i < n;
            const int ic = valid ? i : n - 1, j = ic + w;
            const float2 *pL = C + ic * S1 + 1 + u.a0, *pR = C + j * S + ic + 1 + u.a0;
            const float *tL = Ct + ic * S1 + 2 + u.a0;
            const float2 *iL = I + j * S + ic + 1 + u.a0, *iR = I + ic * S1 + 1 + u.a0;
            const float *tR = Ct + j * S + ic + 1 + u.a0;
            float vxr = NEG_BIG, vxl = NEG_BIG, vl0 = NEG_BIG, vl1 = NEG_BIG, vr0 = NEG_BIG, vr1 = NEG_BIG;
            int axr = 255, axl = 255, al0 = 255, al1 = 255, ar0 = 255, ar1 = 255;
#pragma unroll 4
            for (int k = 0; k < u.cnt; ++k) {
                const int a = u.a0 + k;
                const float2 t = __fadd2_rn(pL[k], pR[k]);  // (XR term, XL term), split a
                amax(vxr, axr, t.x, a);
                amax(vxl, axl, t.y, a);
                const float l3 = tL[k], r4 = tR[k];
                const float2 i3 = iL[k], i4 = iR[k];
                amax(vl0, al0, __fadd_rn(l3, i3.x), a + 1);  // step 3, split a + 1
                amax(vl1, al1, __fadd_rn(l3, i3.y), a + 1);
                amax(vr0, ar0, __fadd_rn(i4.x, r4), a);      // step 4, split a
                amax(vr1, ar1, __fadd_rn(i4.y, r4), a);
            }
            for (int o = 16; o >= u.spw; o >>= 1) {
                amerge(vxr, axr, o); amerge(vxl, axl, o);
                amerge(vl0, al0, o); amerge(vl1, al1, o); amerge(vr0, ar0, o); amerge(vr1, ar1, o);
            }
            if (valid && u.q == 0) {
                const float2 t = __fadd2_rn(C[ic * S + ic + w], C[j * S + j]);  // X term a = w - 1: the largest split
                amax(vxr, axr, t.x, w - 1);
                amax(vxl, axl, t.y, w - 1);
                float2 *pil = I + j * S + ic, *pir = I + ic * S + j;
                const float2 al = *pil, ar = *pir;
                const float2 il = make_float2(__fadd_rn(vxl, al.x), __fadd_rn(vxl, al.y));
                const float2 ir = make_float2(__fadd_rn(vxr, ar.x), __fadd_rn(vxr, ar.y));
                const float l0 = Ct[ic * S1 + 1], r0 = Ct[j * S + j];  // CL(i,i).NO, CR(j,j).NO
                float t0 = __fadd_rn(l0, il.x), t1 = __fadd_rn(l0, il.y);   // step 3, split 0: first index
                if (t0 >= vl0) { vl0 = t0; al0 = 0; }
                if (t1 >= vl1) { vl1 = t1; al1 = 0; }
                t0 = __fadd_rn(ir.x, r0); t1 = __fadd_rn(ir.y, r0);         // step 4, split w - 1: last index
                amax(vr0, ar0, t0, w - 1);
                amax(vr1, ar1, t1, w - 1);
                if (ic == 0 && w != len) { vr0 = p.mask_zero; vr1 = p.mask_zero; }
                *pil = il; *pir = ir;
                BX[j * S + ic] = (uint8_t)axl;
                BX[ic * S + j] = (uint8_t)axr;
                C[j * S + ic] = make_float2(vl1, vl0);
                C[ic * S + j + 1] = make_float2(vr0, vr1);
                Ct[ic * S + j + 1] = vl1;
                Ct[j * S + ic] = vr1;
                BC[j * S + ic] = make_uchar2((uint8_t)al0, (uint8_t)al1);
                BC[ic * S + j + 1] = make_uchar2((uint8_t)ar0, (uint8_t)ar1);
            }
        }
        blk_sync<NT>();
    }
    if (tid == 0) p.best[b] = C[len + 1].y;

    // back-trace: breadth-first over the derivation, one warp, two children per expanded item
    if (tid < 32 && (p.heads || p.arcs || p.vgdec)) {
        const int qcap = 2 * Nb + 2;
        int *cur = queue, *nxt = queue + qcap;
        int ncur = 1;
        if (lane == 0) cur[0] = mk_item(0, 1, 0, len);
        __syncwarp();
        while (ncur > 0) {
            int nnext = 0;
            for (int base = 0; base < ncur; base += 32) {
                const int idx = base + lane;
                int c1 = -1, c2 = -1;
                if (idx < ncur) {
                    const int it = cur[idx];
                    const int kind = it & 3, v = (it >> 2) & 1, lo = (it >> 3) & 511, hi = it >> 12;
                    if (kind < 2 && hi == lo) {  // STOP decision of position lo; kind 0 = right side
                        if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + lo * 8 + (kind == 0 ? 4 : 0) + v * 2 + 1], 1.f);
                    } else if (kind == 0) {  // CR(lo,hi,v) -> IR(lo,r,v) + CR(r,hi,NO), r = lo+1+bp
                        const uchar2 bp = BC[lo * S + hi + 1];
                        const int r = lo + 1 + (int)(v ? bp.y : bp.x);
                        c1 = mk_item(2, v, lo, r); c2 = mk_item(0, 1, r, hi);
                    } else if (kind == 1) {  // CL(hi,lo,v) -> CL(r,lo,NO) + IL(hi,r,v), r = lo+bp
                        const uchar2 bp = BC[hi * S + lo];
                        const int r = lo + (int)(v ? bp.y : bp.x);
                        c1 = mk_item(1, 1, lo, r); c2 = mk_item(3, v, r, hi);
                    } else if (kind == 2) {  // IR: arc lo -> hi; XR -> CR(lo,r,HAS) + CL(hi,r+1,NO)
                        const int r = lo + (int)BX[lo * S + hi];
                        c1 = mk_item(0, 0, lo, r); c2 = mk_item(1, 1, r + 1, hi);
                        if (p.heads) p.heads[(size_t)b * N + hi] = lo;
                        if (p.arcs) p.arcs[(((size_t)b * N + lo) * N + hi) * 2 + v] = 1.f;
                        if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + lo * 8 + 4 + v * 2 + 0], 1.f);
                    } else {  // IL: arc hi -> lo; XL -> CR(lo,r,NO) + CL(hi,r+1,HAS)
                        const int r = lo + (int)BX[hi * S + lo];
                        c1 = mk_item(0, 1, lo, r); c2 = mk_item(1, 0, r + 1, hi);
                        if (p.heads) p.heads[(size_t)b * N + lo] = hi;
                        if (p.arcs) p.arcs[(((size_t)b * N + hi) * N + lo) * 2 + v] = 1.f;
                        if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + hi * 8 + 0 + v * 2 + 0], 1.f);
                    }
                }
                const unsigned has = __ballot_sync(0xffffffffu, c1 >= 0);
                if (c1 >= 0) {
                    const int pos = nnext + 2 * __popc(has & ((1u << lane) - 1u));
                    nxt[pos] = c1; nxt[pos + 1] = c2;
                }
                nnext += 2 * __popc(has);
            }
            __syncwarp();
            int *t = cur; cur = nxt; nxt = t;
            ncur = nnext;
        }
    }
    blk_sync<NT>();
}

__host__ __device__ inline size_t log_bytes(int cap) {
    const size_t E = (size_t)cap * row_stride(cap);
    return E * 40 + (size_t)cap * 32 + (size_t)(cap + 4) * 4;  // + per-word offsets
}
__host__ __device__ inline size_t max_bytes(int cap) {
    const size_t E = (size_t)cap * row_stride(cap);
    return E * 20 + (size_t)cap * 32 + (size_t)(2 * (2 * cap + 2)) * 4 + E * 2 + E + 16;
}
size_t gather_bytes(int cap, int passes) {
    size_t s = 0;
    if (passes & 1) s = log_bytes(cap);
    if (passes & 2) { const size_t m = max_bytes(cap); s = m > s ? m : s; }
    return (s + 15) & ~(size_t)15;
}

__device__ __forceinline__ int clamp_len(const DmvArgs &p, int b) {
    const int len = (int)p.lengths[b];
    return len < 0 ? 0 : (len > p.N - 1 ? p.N - 1 : len);
}

// One CTA per sentence (NT >= 64) or WPC warps = WPC sentences per CTA (NT == 32); persistent, strided over the
// (sentence, semiring) work items of the launch's length bucket.
template <int NT, int WPC, int MINB>
__global__ void __launch_bounds__(NT * WPC, MINB) dmv_gather_kernel(DmvArgs p, int slice_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int sub = NT == 32 ? (int)(threadIdx.x >> 5) : 0;
    const int tid = NT == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    unsigned char *slice = smem_raw + (size_t)sub * slice_bytes;
    // work items = (sentence, semiring), log items first (they cost about twice a max item)
    const int total = p.B * p.npass, step = gridDim.x * WPC;
    for (int item = blockIdx.x * WPC + sub; item < total; item += step) {
        int b, which;
        if (p.npass == 2) { which = item >= p.B; b = which ? item - p.B : item; }
        else { which = p.first_pass; b = item; }
        const int len = clamp_len(p, b);
        if (len + 1 < p.nb_lo || len + 1 > p.nb_hi) continue;
        if (which == 0) {
            if (p.log_domain || !lin_pass<NT>(p, b, len, slice, tid)) log_pass<NT>(p, b, len, slice, tid);
        } else {
            max_pass<NT>(p, b, len, slice, tid);
        }
    }
}

}  // namespace

bool dmv_gather_fits(int cap, int passes, int smem_optin) { return cap <= 256 && gather_bytes(cap, passes) <= (size_t)smem_optin; }

cudaError_t launch_dmv_gather(DmvArgs a, int passes, int cap, int threads, int sm_count, cudaStream_t st) {
    const int total = a.B * a.npass;
    const size_t slice = gather_bytes(cap, passes);
    auto go = [&](auto kern, int nt, int wpc) -> cudaError_t {
        const size_t smem = slice * wpc;
        struct Cached { const void *fn; size_t smem; int occ, dev; };
        static thread_local Cached cache[16];
        static thread_local int ncache = 0;
        int occ = 0, dev = 0;
        cudaGetDevice(&dev);
        for (int k = 0; k < ncache; ++k)
            if (cache[k].fn == (const void *)kern && cache[k].smem == smem && cache[k].dev == dev) occ = cache[k].occ;
        if (occ == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt * wpc, smem);
            if (e != cudaSuccess) return e;
            if (occ < 1) occ = 1;
            int slot = -1;
            for (int k = 0; k < ncache; ++k) if (cache[k].fn == (const void *)kern && cache[k].dev == dev) slot = k;
            if (slot < 0 && ncache < 16) slot = ncache++;
            if (slot >= 0) cache[slot] = Cached{(const void *)kern, smem, occ, dev};
        }
        int grid = sm_count * occ;
        const int need = (total + wpc - 1) / wpc;
        if (grid > need) grid = need;
        kern<<<grid, nt * wpc, smem, st>>>(a, (int)slice);
        return cudaGetLastError();
    };
    if (threads <= 32) return go(dmv_gather_kernel<32, 4, 3>, 32, 4);
    if (threads <= 64) return go(dmv_gather_kernel<64, 1, 6>, 64, 1);
    if (threads <= 128) return go(dmv_gather_kernel<128, 1, 3>, 128, 1);
    return go(dmv_gather_kernel<256, 1, 1>, 256, 1);
}

}  // namespace vlgae
