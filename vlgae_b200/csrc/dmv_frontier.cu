// dmv_frontier.cu -- DMV chart DP, "frontier" schedule: one CTA per sentence, one thread per (target cell, new terms).
//
// Same operator as dmv_kernels.cu (reference /root/reference/src/model/torch_struct/dmv.py:19-66 and the autograd
// marginals / argmax of helpers.py:118-154), different schedule.  The role-split kernel there evaluates a width-w item
// as "gather w terms, reduce across lanes": two shuffle trees, two barriers and ~350 instructions per warp on the
// critical path of every width, 2 (len) widths deep -- latency-bound at ~1.7k clk per width.
//
// Here every target cell keeps a running (max, sum) -- or (best, first arg-max) -- and a term is folded in during the
// phase in which its LATER operand becomes final.  With operand widths (a, b) a term of steps 1/2 (a + b = w - 1) is
// ready after the complete items of width max(a, b); a term of steps 3/4 (complete a, incomplete w - a) after the
// incomplete items of width w - a (if a < w - a) or the complete items of width a.  So phase
//   A(s): incomplete items of width s are final  -> fold them into complete targets of width s .. 2s-1, finalise C(s)
//   B(s): complete items of width s are final    -> fold them into incomplete targets of width s+1 .. 2s+1 and
//                                                   complete targets of width s+1 .. 2s, finalise I(s+1)
// gives every target at most TWO new terms per phase: no cross-lane reduction, no serial loop over split points; the
// critical path of a phase is load -> add -> max -> ex2 -> fma -> store + one barrier (~150-250 clk), and the work of
// a phase ((Nb - w) cells over a band of widths, <= ~600 for 40 words) is one task per thread.
// The reverse sweep is term-parallel by construction: in phase A'(w) / B'(w) the (Nb - w) * w terms of the width-w
// parents each push one product into their two operands; within a phase every accumulator word has a single writer
// (row owner / column owner / distinct words per item kind), so no atomics.
//
// Shared memory per cell (diagonal-major index as in dmv_kernels.cu), log pass 80 B:
//   C4 = (CL.HAS, CL.NO, CR.HAS, CR.NO)   I4 = (IL.HAS, IL.NO, IR.HAS, IR.NO), pre-loaded with attach + dec[GO]
//   A0 = inside: running (m, s) of XL, XR      -> after finalisation / outside: (XL, XR, -, -)
//   A1 = inside: running (m, s) of CL.HAS, CL.NO -> outside: gradient of I4
//   A2 = inside: running (m, s) of CR.HAS, CR.NO -> outside: gradient of C4
// max pass 62 B: C4, I4, VX = best (XL, XR), VC = best (CL.HAS, CL.NO, CR.HAS, CR.NO), 6 arg-max bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dmv_kernels.cuh"

namespace vlgae {

namespace {

constexpr float NEG_BIG = -3.0e38f;  // finite stand-in for -inf
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fexp(float x) { return ex2(x * LOG2E); }
__device__ __forceinline__ float flog(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * LN2;
}

__host__ __device__ inline int ncells(int Nb) { return Nb * (Nb + 1) / 2; }
__device__ __forceinline__ int dbase(int d, int Nb) { return d * Nb - ((d * (d - 1)) >> 1); }
__device__ __forceinline__ int cidx(int lo, int d, int Nb) { return dbase(d, Nb) + lo; }

// running logsumexp (m, s): value = m + log s
__device__ __forceinline__ void lse1(float &m, float &s, float t) {
    const float d = t - m;
    const float e = ex2(-fabsf(d) * LOG2E);
    if (d > 0.f) { s = fmaf(s, e, 1.f); m = t; }
    else s += e;
}
__device__ __forceinline__ void lse2(float &m, float &s, float t1, float t2) {
    const float mm = fmaxf(m, fmaxf(t1, t2));
    s = fmaf(s, ex2((m - mm) * LOG2E), ex2((t1 - mm) * LOG2E) + ex2((t2 - mm) * LOG2E));
    m = mm;
}
__device__ __forceinline__ float lse_fin(float m, float s) { return m + flog(s); }

// running first-arg-max: the reference takes torch.max over the split points in index order (semirings.py:200-202),
// i.e. the SMALLEST index among equal values; terms arrive out of order here, so ties compare the index.
__device__ __forceinline__ void amax1(float &v, int &a, float t, int idx) {
    if (t > v || (t == v && idx < a)) { v = t; a = idx; }
}

struct LogChart {
    float4 *C4, *I4, *A0, *A1, *A2;
};
struct MaxChart {
    float4 *C4, *I4, *VC;
    float2 *VX;
    uint8_t *bp;  // 6 bytes per cell: XL, XR, CL[HAS], CL[NO], CR[HAS], CR[NO]  (first maximal split)
};

__device__ __forceinline__ int clamp_len(const DmvArgs &p, int b) {
    int len = (int)p.lengths[b];
    return len < 0 ? 0 : (len > p.N - 1 ? p.N - 1 : len);
}

// stage dec, width-0 complete items (STOP decisions, dmv.py:39-40) and the arc scores attach + dec[GO]
// (formed first in fp32, exactly as dmv.py:36-37 does).  dec index = dir*4 + val*2 + decision.
template <int NT>
__device__ __forceinline__ void stage_inputs(const DmvArgs &p, int b, int Nb, float *sdec, uint16_t *cw, float4 *C4, float4 *I4) {
    const int tid = threadIdx.x, N = p.N;
    const float *dec = p.dec + (size_t)b * N * 8;
    const float *attach = p.attach + (size_t)b * N * N * 2;
    #pragma unroll 1
    for (int t = tid; t < Nb * 8; t += NT) sdec[t] = dec[t];
    #pragma unroll 1
    for (int d = tid; d < Nb; d += NT) {  // cell -> (width, left end)
        const int base = dbase(d, Nb);
        for (int lo = 0; lo < Nb - d; ++lo) cw[base + lo] = (uint16_t)((d << 8) | lo);
    }
    __syncthreads();
    #pragma unroll 1
    for (int i = tid; i < Nb; i += NT) C4[i] = make_float4(sdec[i * 8 + 1], sdec[i * 8 + 3], sdec[i * 8 + 5], sdec[i * 8 + 7]);
    const int nc = ncells(Nb);
    #pragma unroll 1
    for (int c = Nb + tid; c < nc; c += NT) {
        const int w = cw[c] >> 8, i = cw[c] & 255, j = i + w;
        const float2 al = *reinterpret_cast<const float2 *>(attach + ((size_t)j * N + i) * 2);  // arc j -> i
        const float2 ar = *reinterpret_cast<const float2 *>(attach + ((size_t)i * N + j) * 2);  // arc i -> j
        I4[c] = make_float4(__fadd_rn(al.x, sdec[j * 8 + 0]), __fadd_rn(al.y, sdec[j * 8 + 2]),
                            __fadd_rn(ar.x, sdec[i * 8 + 4]), __fadd_rn(ar.y, sdec[i * 8 + 6]));
    }
}

// ---------------------------------------------------------------------------------------------
// register-resident variants: thread t owns the target cells Nb + t + k NT (k < CPT) for the whole sweep and keeps
// their running state in registers -- no cell decode, no accumulator load / store, no task loop per phase.
// ---------------------------------------------------------------------------------------------
template <int NT, int CPT>
__device__ __forceinline__ void inside_reg(const LogChart &c, const uint16_t *cw, int Nb, int len, float mask_zero) {
    const int tid = threadIdx.x, nc = ncells(Nb);
    int ow[CPT], oi[CPT];
    float ax[CPT][4], al[CPT][4], ar[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cc = Nb + tid + k * NT;
        const int e = cc < nc ? (int)cw[cc] : 0;
        ow[k] = e >> 8; oi[k] = e & 255;  // width 0 = no cell: never inside a band
#pragma unroll
        for (int q = 0; q < 4; ++q) { ax[k][q] = (q & 1) ? 0.f : NEG_BIG; al[k][q] = ax[k][q]; ar[k][q] = ax[k][q]; }
    }
#pragma unroll 1
    for (int s = 0; s <= len; ++s) {
        const int Ds = dbase(s, Nb);
        if (s >= 1) {
            // phase A(s): incomplete items of width s are final
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int w = ow[k], i = oi[k];
                if (w >= s && w <= 2 * s - 1) {
                    const int Dd = dbase(w - s, Nb), j = i + w;
                    const float l3 = c.C4[Dd + i].y;          // CL[i, j-s].NO
                    const float4 i3 = c.I4[Ds + j - s];        // IL[j-s, j]
                    const float4 i4 = c.I4[Ds + i];            // IR[i, i+s]
                    const float r4 = c.C4[Dd + i + s].w;       // CR[i+s, j].NO
                    lse1(al[k][0], al[k][1], l3 + i3.x);
                    lse1(al[k][2], al[k][3], l3 + i3.y);
                    lse1(ar[k][0], ar[k][1], i4.z + r4);
                    lse1(ar[k][2], ar[k][3], i4.w + r4);
                    if (w == s) {
                        float4 v = make_float4(lse_fin(al[k][0], al[k][1]), lse_fin(al[k][2], al[k][3]),
                                               lse_fin(ar[k][0], ar[k][1]), lse_fin(ar[k][2], ar[k][3]));
                        if (i == 0 && w != len) { v.z = mask_zero; v.w = mask_zero; }  // single-root mask, dmv.py:63
                        c.C4[Nb + tid + k * NT] = v;
                    }
                }
            }
            __syncthreads();
        }
        if (s == len) break;
        // phase B(s): complete items of width s are final
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int w = ow[k], i = oi[k];
            if (w >= s + 1 && w <= 2 * s + 1) {
                const int De = dbase(w - 1 - s, Nb), j = i + w;
                const float4 la = c.C4[Ds + i], ra = c.C4[De + i + s + 1];
                if (w - 1 - s != s) {
                    const float4 lb = c.C4[De + i], rb = c.C4[Ds + j - s];
                    lse2(ax[k][0], ax[k][1], la.w + ra.x, lb.w + rb.x);
                    lse2(ax[k][2], ax[k][3], la.z + ra.y, lb.z + rb.y);
                } else {
                    lse1(ax[k][0], ax[k][1], la.w + ra.x);
                    lse1(ax[k][2], ax[k][3], la.z + ra.y);
                }
                if (w == s + 1) {
                    const int cc = Nb + tid + k * NT;
                    const float xl = lse_fin(ax[k][0], ax[k][1]), xr = lse_fin(ax[k][2], ax[k][3]);
                    const float4 arc = c.I4[cc];
                    c.I4[cc] = make_float4(xl + arc.x, xl + arc.y, xr + arc.z, xr + arc.w);
                    c.A0[cc] = make_float4(xl, xr, 0.f, 0.f);
                }
                if (w <= 2 * s) {
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = la.y;                     // CL[i, i+s].NO
                    const float4 i3 = c.I4[Dd + i + s];        // IL[i+s, j]
                    const float4 i4 = c.I4[Dd + i];            // IR[i, j-s]
                    const float r4 = c.C4[Ds + j - s].w;       // CR[j-s, j].NO
                    lse1(al[k][0], al[k][1], l3 + i3.x);
                    lse1(al[k][2], al[k][3], l3 + i3.y);
                    lse1(ar[k][0], ar[k][1], i4.z + r4);
                    lse1(ar[k][2], ar[k][3], i4.w + r4);
                }
            }
        }
        __syncthreads();
    }
}

template <int NT, int CPT>
__device__ __forceinline__ void viterbi_reg(const MaxChart &c, const uint16_t *cw, int Nb, int len, float mask_zero) {
    const int tid = threadIdx.x, nc = ncells(Nb);
    int ow[CPT], oi[CPT];
    float vx[CPT][2], vc[CPT][4];
    int bx[CPT][2], bc[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cc = Nb + tid + k * NT;
        const int e = cc < nc ? (int)cw[cc] : 0;
        ow[k] = e >> 8; oi[k] = e & 255;
        vx[k][0] = vx[k][1] = NEG_BIG; bx[k][0] = bx[k][1] = 255;
#pragma unroll
        for (int q = 0; q < 4; ++q) { vc[k][q] = NEG_BIG; bc[k][q] = 255; }
    }
#pragma unroll 1
    for (int s = 0; s <= len; ++s) {
        const int Ds = dbase(s, Nb);
        if (s >= 1) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int w = ow[k], i = oi[k];
                if (w >= s && w <= 2 * s - 1) {
                    const int Dd = dbase(w - s, Nb), j = i + w;
                    const float l3 = c.C4[Dd + i].y;
                    const float4 i3 = c.I4[Ds + j - s];
                    const float4 i4 = c.I4[Ds + i];
                    const float r4 = c.C4[Dd + i + s].w;
                    amax1(vc[k][0], bc[k][0], __fadd_rn(l3, i3.x), w - s);   // CL split r - i, r = j - s
                    amax1(vc[k][1], bc[k][1], __fadd_rn(l3, i3.y), w - s);
                    amax1(vc[k][2], bc[k][2], __fadd_rn(i4.z, r4), s - 1);   // CR split r - i - 1, r = i + s
                    amax1(vc[k][3], bc[k][3], __fadd_rn(i4.w, r4), s - 1);
                    if (w == s) {
                        const int cc = Nb + tid + k * NT;
                        float4 v = make_float4(vc[k][0], vc[k][1], vc[k][2], vc[k][3]);
                        if (i == 0 && w != len) { v.z = mask_zero; v.w = mask_zero; }
                        c.C4[cc] = v;
                        uint8_t *bp = c.bp + cc * 6;
                        bp[2] = (uint8_t)bc[k][0]; bp[3] = (uint8_t)bc[k][1]; bp[4] = (uint8_t)bc[k][2]; bp[5] = (uint8_t)bc[k][3];
                    }
                }
            }
            __syncthreads();
        }
        if (s == len) break;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int w = ow[k], i = oi[k];
            if (w >= s + 1 && w <= 2 * s + 1) {
                const int De = dbase(w - 1 - s, Nb), j = i + w;
                const float4 la = c.C4[Ds + i], ra = c.C4[De + i + s + 1];
                amax1(vx[k][0], bx[k][0], __fadd_rn(la.w, ra.x), s);
                amax1(vx[k][1], bx[k][1], __fadd_rn(la.z, ra.y), s);
                if (w - 1 - s != s) {
                    const float4 lb = c.C4[De + i], rb = c.C4[Ds + j - s];
                    amax1(vx[k][0], bx[k][0], __fadd_rn(lb.w, rb.x), w - 1 - s);
                    amax1(vx[k][1], bx[k][1], __fadd_rn(lb.z, rb.y), w - 1 - s);
                }
                if (w == s + 1) {
                    const int cc = Nb + tid + k * NT;
                    const float4 arc = c.I4[cc];
                    c.I4[cc] = make_float4(__fadd_rn(vx[k][0], arc.x), __fadd_rn(vx[k][0], arc.y),
                                           __fadd_rn(vx[k][1], arc.z), __fadd_rn(vx[k][1], arc.w));
                    uint8_t *bp = c.bp + cc * 6;
                    bp[0] = (uint8_t)bx[k][0]; bp[1] = (uint8_t)bx[k][1];
                }
                if (w <= 2 * s) {
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = la.y;
                    const float4 i3 = c.I4[Dd + i + s];
                    const float4 i4 = c.I4[Dd + i];
                    const float r4 = c.C4[Ds + j - s].w;
                    amax1(vc[k][0], bc[k][0], __fadd_rn(l3, i3.x), s);
                    amax1(vc[k][1], bc[k][1], __fadd_rn(l3, i3.y), s);
                    amax1(vc[k][2], bc[k][2], __fadd_rn(i4.z, r4), w - s - 1);
                    amax1(vc[k][3], bc[k][3], __fadd_rn(i4.w, r4), w - s - 1);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// log semiring: inside + outside for one sentence
// ---------------------------------------------------------------------------------------------
template <int NT, int CPT>
__device__ void log_pass(const DmvArgs &p, int b, unsigned char *mem) {
    const int tid = threadIdx.x, N = p.N;
    const int len = clamp_len(p, b), Nb = len + 1, nc = ncells(Nb);
    float *sdec = reinterpret_cast<float *>(mem);
    LogChart c;
    c.C4 = reinterpret_cast<float4 *>(mem + (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15));
    c.I4 = c.C4 + nc; c.A0 = c.I4 + nc; c.A1 = c.A0 + nc; c.A2 = c.A1 + nc;
    uint16_t *cw = reinterpret_cast<uint16_t *>(c.A2 + nc);
    const bool want_grad = (p.gdec != nullptr) || (p.gattach != nullptr);
    const bool prof = p.prof && b == 0 && tid == 0;
    long long t0c = 0;
    if (prof) t0c = clock64();

    stage_inputs<NT>(p, b, Nb, sdec, cw, c.C4, c.I4);
    {
        const float4 init = make_float4(NEG_BIG, 0.f, NEG_BIG, 0.f);
        #pragma unroll 1
        for (int t = Nb + tid; t < nc; t += NT) { c.A0[t] = init; c.A1[t] = init; c.A2[t] = init; }
    }
    __syncthreads();
    if (prof) p.prof[0] = clock64() - t0c;

    // ---------------- inside ----------------
    if (CPT > 0 && nc - Nb <= CPT * NT) {
        inside_reg<NT, (CPT > 0 ? CPT : 1)>(c, cw, Nb, len, p.mask_zero);
    } else {
    #pragma unroll 1
    for (int s = 0; s <= len; ++s) {
        if (s >= 1) {
            // phase A(s): incomplete items of width s are final
            const int whi = min(2 * s - 1, len);
            const int c0 = dbase(s, Nb), c1 = dbase(whi + 1, Nb);
            #pragma unroll 1
            for (int cc = c0 + tid; cc < c1; cc += NT) {
                const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                // step 3 (dmv.py:58-59): CL[i,j][v] (+)= CL[i,r].NO + IL[r,j][v], r = j - s
                const float l3 = c.C4[cidx(i, w - s, Nb)].y;
                const float4 i3 = c.I4[cidx(j - s, s, Nb)];
                // step 4 (dmv.py:61-62): CR[i,j][v] (+)= IR[i,r][v] + CR[r,j].NO, r = i + s
                const float4 i4 = c.I4[cidx(i, s, Nb)];
                const float r4 = c.C4[cidx(i + s, w - s, Nb)].w;
                float4 a1 = c.A1[cc], a2 = c.A2[cc];
                lse1(a1.x, a1.y, l3 + i3.x);
                lse1(a1.z, a1.w, l3 + i3.y);
                lse1(a2.x, a2.y, i4.z + r4);
                lse1(a2.z, a2.w, i4.w + r4);
                if (w == s) {
                    float4 v = make_float4(lse_fin(a1.x, a1.y), lse_fin(a1.z, a1.w), lse_fin(a2.x, a2.y), lse_fin(a2.z, a2.w));
                    if (i == 0 && w != len) { v.z = p.mask_zero; v.w = p.mask_zero; }  // single-root mask, dmv.py:63
                    c.C4[cc] = v;
                } else {
                    c.A1[cc] = a1; c.A2[cc] = a2;
                }
            }
            __syncthreads();
        }
        if (s == len) break;
        // phase B(s): complete items of width s are final
        {
            const int ihi = min(2 * s + 1, len), chi = min(2 * s, len);
            const int i0 = dbase(s + 1, Nb), nI = dbase(ihi + 1, Nb) - i0;
            const int nC = chi >= s + 1 ? dbase(chi + 1, Nb) - i0 : 0;
            #pragma unroll 1
            for (int t = tid; t < nI + nC; t += NT) {
                if (t < nI) {
                    const int cc = i0 + t;
                    const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                    // steps 1, 2 (dmv.py:50-56): XL (+)= CR[i,r].NO + CL[r+1,j].HAS, XR (+)= CR[i,r].HAS + CL[r+1,j].NO
                    const float4 la = c.C4[cidx(i, s, Nb)], ra = c.C4[cidx(i + s + 1, w - 1 - s, Nb)];
                    float4 a0 = c.A0[cc];
                    if (w - 1 - s != s) {
                        const float4 lb = c.C4[cidx(i, w - 1 - s, Nb)], rb = c.C4[cidx(j - s, s, Nb)];
                        lse2(a0.x, a0.y, la.w + ra.x, lb.w + rb.x);
                        lse2(a0.z, a0.w, la.z + ra.y, lb.z + rb.y);
                    } else {
                        lse1(a0.x, a0.y, la.w + ra.x);
                        lse1(a0.z, a0.w, la.z + ra.y);
                    }
                    if (w == s + 1) {
                        const float xl = lse_fin(a0.x, a0.y), xr = lse_fin(a0.z, a0.w);
                        const float4 arc = c.I4[cc];
                        c.I4[cc] = make_float4(xl + arc.x, xl + arc.y, xr + arc.z, xr + arc.w);
                        c.A0[cc] = make_float4(xl, xr, 0.f, 0.f);
                    } else {
                        c.A0[cc] = a0;
                    }
                } else {
                    const int cc = i0 + (t - nI);
                    const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                    const float l3 = c.C4[cidx(i, s, Nb)].y;               // CL[i, i+s].NO
                    const float4 i3 = c.I4[cidx(i + s, w - s, Nb)];         // IL[i+s, j]
                    const float4 i4 = c.I4[cidx(i, w - s, Nb)];             // IR[i, j-s]
                    const float r4 = c.C4[cidx(j - s, s, Nb)].w;            // CR[j-s, j].NO
                    float4 a1 = c.A1[cc], a2 = c.A2[cc];
                    lse1(a1.x, a1.y, l3 + i3.x);
                    lse1(a1.z, a1.w, l3 + i3.y);
                    lse1(a2.x, a2.y, i4.z + r4);
                    lse1(a2.z, a2.w, i4.w + r4);
                    c.A1[cc] = a1; c.A2[cc] = a2;
                }
            }
            __syncthreads();
        }
    }
    }
    if (prof) p.prof[1] = clock64() - t0c;
    if (tid == 0) p.Z[b] = c.C4[cidx(0, len, Nb)].w;  // dmv.py:65
    if (!want_grad) { __syncthreads(); return; }

    // ---------------- outside (explicit reverse sweep) ----------------
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        #pragma unroll 1
        for (int t = tid; t < nc; t += NT) { c.A1[t] = z; c.A2[t] = z; }
        __syncthreads();
        if (tid == 0) c.A2[cidx(0, len, Nb)].w = p.gZ ? p.gZ[b] : 1.f;
        __syncthreads();
    }
    float *gI = reinterpret_cast<float *>(c.A1), *gC = reinterpret_cast<float *>(c.A2);
    #pragma unroll 1
    for (int w = len; w >= 1; --w) {
        const int ntask = (Nb - w) * w;
        const float rw = 1.0f / (float)w;
        const int pb = dbase(w, Nb);
        // phase A'(w): complete parents of width w (steps 3, 4 transposed)
        #pragma unroll 1
        for (int t = tid; t < ntask; t += NT) {
            const int i = (int)(((float)t + 0.5f) * rw), a = t - i * w, j = i + w;
            const float4 gg = c.A2[pb + i], out = c.C4[pb + i];
            {
                const int cl = cidx(i, a, Nb), cr = cidx(i + a, w - a, Nb);
                const float lv = c.C4[cl].y;
                const float4 rv = c.I4[cr];
                const float p0 = gg.x * fexp(lv + rv.x - out.x), p1 = gg.y * fexp(lv + rv.y - out.y);
                gC[cl * 4 + 1] += p0 + p1;
                gI[cr * 4 + 0] += p0;
                gI[cr * 4 + 1] += p1;
            }
            if (!(i == 0 && w != len)) {  // the mask overwrote CR[0][w]: no gradient passes through it
                const int cl = cidx(i, a + 1, Nb), cr = cidx(i + a + 1, w - 1 - a, Nb);
                const float4 lv = c.I4[cl];
                const float rv = c.C4[cr].w;
                const float q0 = gg.z * fexp(lv.z + rv - out.z), q1 = gg.w * fexp(lv.w + rv - out.w);
                gI[cl * 4 + 2] += q0;
                gI[cl * 4 + 3] += q1;
                gC[cr * 4 + 3] += q0 + q1;
            }
            (void)j;
        }
        __syncthreads();
        // phase B'(w): incomplete parents of width w (steps 1, 2 transposed)
        #pragma unroll 1
        for (int t = tid; t < ntask; t += NT) {
            const int i = (int)(((float)t + 0.5f) * rw), a = t - i * w;
            const float4 gi = c.A1[pb + i];
            const float4 x = c.A0[pb + i];
            const float gl = gi.x + gi.y, gr = gi.z + gi.w;
            const int cl = cidx(i, a, Nb), cr = cidx(i + a + 1, w - 1 - a, Nb);
            const float4 lv = c.C4[cl], rv = c.C4[cr];
            const float pl = gl * fexp(lv.w + rv.x - x.x), pr = gr * fexp(lv.z + rv.y - x.y);
            gC[cl * 4 + 3] += pl;
            gC[cl * 4 + 2] += pr;
            gC[cr * 4 + 0] += pl;
            gC[cr * 4 + 1] += pr;
        }
        __syncthreads();
    }
    if (prof) p.prof[2] = clock64() - t0c;

    // ---------------- outputs ----------------
    if (p.gattach) {
        float2 *ga = reinterpret_cast<float2 *>(p.gattach + (size_t)b * N * N * 2);
        #pragma unroll 1
        for (int t = tid; t < N * N; t += NT) {
            const int h = t / N, ch = t - h * N;
            float2 v = make_float2(0.f, 0.f);
            if (h < Nb && ch < Nb && h != ch) {
                const float4 g = ch < h ? c.A1[cidx(ch, h - ch, Nb)] : c.A1[cidx(h, ch - h, Nb)];
                v = ch < h ? make_float2(g.x, g.y) : make_float2(g.z, g.w);
            }
            ga[t] = v;
        }
    }
    if (p.gdec) {
        float *gd = p.gdec + (size_t)b * N * 8;
        #pragma unroll 1
        for (int t = tid; t < N * 2; t += NT) {
            const int i = t >> 1, dir = t & 1;
            float2 go = make_float2(0.f, 0.f), stop = make_float2(0.f, 0.f);
            if (i < Nb) {
                const float4 g0 = c.A2[i];
                if (dir == 0) {
                    for (int ch = 0; ch < i; ++ch) { const float4 v = c.A1[cidx(ch, i - ch, Nb)]; go.x += v.x; go.y += v.y; }
                    stop = make_float2(g0.x, g0.y);
                } else {
                    for (int d = 1; d < Nb - i; ++d) { const float4 v = c.A1[cidx(i, d, Nb)]; go.x += v.z; go.y += v.w; }
                    stop = make_float2(g0.z, g0.w);
                }
            }
            *reinterpret_cast<float4 *>(gd + i * 8 + dir * 4) = make_float4(go.x, stop.x, go.y, stop.y);  // [dir][val][decision]
        }
    }
    __syncthreads();
    if (prof) p.prof[3] = clock64() - t0c;
}

// ---------------------------------------------------------------------------------------------
// max semiring: Viterbi chart with first-max back-pointers + breadth-first back-trace
// ---------------------------------------------------------------------------------------------
// items of the back-trace: kind (0 CR, 1 CL, 2 IR, 3 IL) | v << 2 | lo << 3 | hi << 12
__device__ __forceinline__ int mk_item(int kind, int v, int lo, int hi) { return kind | (v << 2) | (lo << 3) | (hi << 12); }

template <int NT, int CPT>
__device__ void max_pass(const DmvArgs &p, int b, unsigned char *mem) {
    const int tid = threadIdx.x, N = p.N;
    const int len = clamp_len(p, b), Nb = len + 1, nc = ncells(Nb);
    float *sdec = reinterpret_cast<float *>(mem);
    MaxChart c;
    c.C4 = reinterpret_cast<float4 *>(mem + (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15));
    c.I4 = c.C4 + nc; c.VC = c.I4 + nc;
    c.VX = reinterpret_cast<float2 *>(c.VC + nc);
    int *queue = reinterpret_cast<int *>(c.VX + nc);  // 2 x (2 Nb + 2) ints
    uint16_t *cw = reinterpret_cast<uint16_t *>(queue + 2 * (2 * Nb + 2));
    c.bp = reinterpret_cast<uint8_t *>(cw + nc + (nc & 1));
    const bool prof = p.prof && b == 0 && tid == 0;
    long long t0c = 0;
    if (prof) t0c = clock64();

    stage_inputs<NT>(p, b, Nb, sdec, cw, c.C4, c.I4);
    #pragma unroll 1
    for (int t = Nb + tid; t < nc; t += NT) {
        c.VC[t] = make_float4(NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG);
        c.VX[t] = make_float2(NEG_BIG, NEG_BIG);
    }
    #pragma unroll 1
    for (int t = tid; t < nc * 6; t += NT) c.bp[t] = 255;
    // outputs that the back-trace only dots with ones are zero-filled up front
    if (p.arcs) {
        float2 *z = reinterpret_cast<float2 *>(p.arcs + (size_t)b * N * N * 2);
        #pragma unroll 1
        for (int t = tid; t < N * N; t += NT) z[t] = make_float2(0.f, 0.f);
    }
    if (p.vgdec) for (int t = tid; t < N * 8; t += NT) p.vgdec[(size_t)b * N * 8 + t] = 0.f;
    if (p.heads) for (int t = tid; t < N; t += NT) p.heads[(size_t)b * N + t] = 0;
    __syncthreads();
    if (prof) p.prof[4] = clock64() - t0c;

    if (CPT > 0 && nc - Nb <= CPT * NT) {
        viterbi_reg<NT, (CPT > 0 ? CPT : 1)>(c, cw, Nb, len, p.mask_zero);
    } else {
    #pragma unroll 1
    for (int s = 0; s <= len; ++s) {
        if (s >= 1) {
            // phase A(s): the new term of CL is split r - i = w - s, of CR split r - i - 1 = s - 1
            const int whi = min(2 * s - 1, len);
            const int c0 = dbase(s, Nb), c1 = dbase(whi + 1, Nb);
            #pragma unroll 1
            for (int cc = c0 + tid; cc < c1; cc += NT) {
                const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                const float l3 = c.C4[cidx(i, w - s, Nb)].y;
                const float4 i3 = c.I4[cidx(j - s, s, Nb)];
                const float4 i4 = c.I4[cidx(i, s, Nb)];
                const float r4 = c.C4[cidx(i + s, w - s, Nb)].w;
                float4 v = c.VC[cc];
                uint8_t *bp = c.bp + cc * 6;
                int a0 = bp[2], a1 = bp[3], a2 = bp[4], a3 = bp[5];
                amax1(v.x, a0, __fadd_rn(l3, i3.x), w - s);
                amax1(v.y, a1, __fadd_rn(l3, i3.y), w - s);
                amax1(v.z, a2, __fadd_rn(i4.z, r4), s - 1);
                amax1(v.w, a3, __fadd_rn(i4.w, r4), s - 1);
                bp[2] = (uint8_t)a0; bp[3] = (uint8_t)a1; bp[4] = (uint8_t)a2; bp[5] = (uint8_t)a3;
                if (w == s) {
                    if (i == 0 && w != len) { v.z = p.mask_zero; v.w = p.mask_zero; }
                    c.C4[cc] = v;
                } else {
                    c.VC[cc] = v;
                }
            }
            __syncthreads();
        }
        if (s == len) break;
        {
            const int ihi = min(2 * s + 1, len), chi = min(2 * s, len);
            const int i0 = dbase(s + 1, Nb), nI = dbase(ihi + 1, Nb) - i0;
            const int nC = chi >= s + 1 ? dbase(chi + 1, Nb) - i0 : 0;
            #pragma unroll 1
            for (int t = tid; t < nI + nC; t += NT) {
                if (t < nI) {
                    const int cc = i0 + t;
                    const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                    const float4 la = c.C4[cidx(i, s, Nb)], ra = c.C4[cidx(i + s + 1, w - 1 - s, Nb)];
                    float2 v = c.VX[cc];
                    uint8_t *bp = c.bp + cc * 6;
                    int a0 = bp[0], a1 = bp[1];
                    amax1(v.x, a0, __fadd_rn(la.w, ra.x), s);  // split r - i = s
                    amax1(v.y, a1, __fadd_rn(la.z, ra.y), s);
                    if (w - 1 - s != s) {
                        const float4 lb = c.C4[cidx(i, w - 1 - s, Nb)], rb = c.C4[cidx(j - s, s, Nb)];
                        amax1(v.x, a0, __fadd_rn(lb.w, rb.x), w - 1 - s);
                        amax1(v.y, a1, __fadd_rn(lb.z, rb.y), w - 1 - s);
                    }
                    bp[0] = (uint8_t)a0; bp[1] = (uint8_t)a1;
                    if (w == s + 1) {
                        const float4 arc = c.I4[cc];
                        c.I4[cc] = make_float4(__fadd_rn(v.x, arc.x), __fadd_rn(v.x, arc.y), __fadd_rn(v.y, arc.z), __fadd_rn(v.y, arc.w));
                    } else {
                        c.VX[cc] = v;
                    }
                } else {
                    const int cc = i0 + (t - nI);
                    const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                    const float l3 = c.C4[cidx(i, s, Nb)].y;
                    const float4 i3 = c.I4[cidx(i + s, w - s, Nb)];
                    const float4 i4 = c.I4[cidx(i, w - s, Nb)];
                    const float r4 = c.C4[cidx(j - s, s, Nb)].w;
                    float4 v = c.VC[cc];
                    uint8_t *bp = c.bp + cc * 6;
                    int a0 = bp[2], a1 = bp[3], a2 = bp[4], a3 = bp[5];
                    amax1(v.x, a0, __fadd_rn(l3, i3.x), s);          // CL split r - i = s
                    amax1(v.y, a1, __fadd_rn(l3, i3.y), s);
                    amax1(v.z, a2, __fadd_rn(i4.z, r4), w - s - 1);  // CR split r - i - 1, r = j - s
                    amax1(v.w, a3, __fadd_rn(i4.w, r4), w - s - 1);
                    bp[2] = (uint8_t)a0; bp[3] = (uint8_t)a1; bp[4] = (uint8_t)a2; bp[5] = (uint8_t)a3;
                    c.VC[cc] = v;
                }
            }
            __syncthreads();
        }
    }
    }
    if (prof) p.prof[5] = clock64() - t0c;
    if (tid == 0) p.best[b] = c.C4[cidx(0, len, Nb)].w;

    // back-trace: breadth-first over the derivation, one warp, two children per expanded item
    if (tid < 32 && (p.heads || p.arcs || p.vgdec)) {
        const int lane = tid;
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
                    const int d = hi - lo;
                    if (kind < 2 && d == 0) {  // STOP decision of position lo; kind 0 = right side
                        if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + lo * 8 + (kind == 0 ? 4 : 0) + v * 2 + 1], 1.f);
                    } else {
                        const uint8_t *bp = c.bp + cidx(lo, d, Nb) * 6;
                        if (kind == 0) {  // CR(lo,hi,v) -> IR(lo,r,v) + CR(r,hi,NO), r = lo+1+bp
                            const int r = lo + 1 + (int)bp[4 + v];
                            c1 = mk_item(2, v, lo, r); c2 = mk_item(0, 1, r, hi);
                        } else if (kind == 1) {  // CL(hi,lo,v) -> CL(r,lo,NO) + IL(hi,r,v), r = lo+bp
                            const int r = lo + (int)bp[2 + v];
                            c1 = mk_item(1, 1, lo, r); c2 = mk_item(3, v, r, hi);
                        } else if (kind == 2) {  // IR: arc lo -> hi; XR -> CR(lo,r,HAS) + CL(hi,r+1,NO)
                            const int r = lo + (int)bp[1];
                            c1 = mk_item(0, 0, lo, r); c2 = mk_item(1, 1, r + 1, hi);
                            if (p.heads) p.heads[(size_t)b * N + hi] = lo;
                            if (p.arcs) p.arcs[(((size_t)b * N + lo) * N + hi) * 2 + v] = 1.f;
                            if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + lo * 8 + 4 + v * 2 + 0], 1.f);
                        } else {  // IL: arc hi -> lo; XL -> CR(lo,r,NO) + CL(hi,r+1,HAS)
                            const int r = lo + (int)bp[0];
                            c1 = mk_item(0, 1, lo, r); c2 = mk_item(1, 0, r + 1, hi);
                            if (p.heads) p.heads[(size_t)b * N + lo] = hi;
                            if (p.arcs) p.arcs[(((size_t)b * N + hi) * N + lo) * 2 + v] = 1.f;
                            if (p.vgdec) atomicAdd(&p.vgdec[(size_t)b * N * 8 + hi * 8 + 0 + v * 2 + 0], 1.f);
                        }
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
    __syncthreads();
    if (prof) p.prof[6] = clock64() - t0c;
}

// ---------------------------------------------------------------------------------------------
// kernel: persistent CTAs stride over (sentence, semiring) work items (same placement rule as dmv_kernels.cu)
// ---------------------------------------------------------------------------------------------
template <int NT, int CPT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : (NT == 256 ? 3 : (NT == 128 ? 6 : (NT == 64 ? 12 : 1)))) dmv_frontier_kernel(DmvArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int total = p.B * p.npass;
    const int nsm = p.nsm;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int item = t;
        if (total <= (int)gridDim.x && t >= nsm) item = total - 1 - (t - nsm);
        int b, which;
        if (p.npass == 2) { which = item >= p.B; b = which ? item - p.B : item; }
        else { which = p.first_pass; b = item; }
        {
            const int len = clamp_len(p, b);
            if (len + 1 < p.nb_lo || len + 1 > p.nb_hi) continue;
        }
        if (which == 0) log_pass<NT, CPT>(p, b, smem_raw);
        else max_pass<NT, CPT>(p, b, smem_raw);
    }
}

size_t frontier_bytes(int cap, int passes) {
    const size_t nc = ncells(cap), dec = ((size_t)cap * 8 * 4 + 15) & ~(size_t)15;
    size_t s = 0;
    if (passes & 1) s = dec + nc * 80 + nc * 2 + 16;
    if (passes & 2) {
        const size_t m = dec + nc * 56 + (size_t)(2 * (2 * cap + 2)) * 4 + (nc + 1) * 2 + nc * 6 + 16;
        s = m > s ? m : s;
    }
    return (s + 255) & ~(size_t)255;
}

}  // namespace

bool dmv_frontier_fits(int cap, int passes, int smem_optin) { return cap <= 256 && frontier_bytes(cap, passes) <= (size_t)smem_optin; }

cudaError_t launch_dmv_frontier(DmvArgs a, int passes, int cap, int threads, bool reg_state, int sm_count, cudaStream_t st) {
    const size_t smem = frontier_bytes(cap, passes);
    const int total = a.B * a.npass;
    a.smem_n = cap;
    auto go = [&](auto kern, int nt) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        int grid = sm_count * occ;
        if (grid > total) grid = total;
        kern<<<grid, nt, smem, st>>>(a);
        return cudaGetLastError();
    };
    // cells per thread of the register-resident sweeps (0 = running state in shared memory, any chart size)
    // running state in registers (thread owns <= 4 target cells for the whole sweep) or in shared memory (any size)
    static const int env_smem_acc = [] { const char *v = getenv("VLGAE_FRONTIER_SMEM_ACC"); return v && *v ? atoi(v) : -1; }();
    if (env_smem_acc >= 0) reg_state = env_smem_acc == 0;
    const int cells = reg_state ? ncells(cap) - cap : (1 << 30);
    if (threads <= 64) return cells <= 128 ? go(dmv_frontier_kernel<64, 2>, 64) : go(dmv_frontier_kernel<64, 0>, 64);
    if (threads <= 128) return cells <= 256 ? go(dmv_frontier_kernel<128, 2>, 128) : go(dmv_frontier_kernel<128, 0>, 128);
    if (threads <= 256) {
        if (cells <= 512) return go(dmv_frontier_kernel<256, 2>, 256);
        return cells <= 1024 ? go(dmv_frontier_kernel<256, 4>, 256) : go(dmv_frontier_kernel<256, 0>, 256);
    }
    if (threads <= 512) {
        if (cells <= 512) return go(dmv_frontier_kernel<512, 1>, 512);
        return cells <= 1024 ? go(dmv_frontier_kernel<512, 2>, 512) : go(dmv_frontier_kernel<512, 0>, 512);
    }
    return cells <= 1024 ? go(dmv_frontier_kernel<1024, 1>, 1024) : go(dmv_frontier_kernel<1024, 0>, 1024);
}

}  // namespace vlgae
