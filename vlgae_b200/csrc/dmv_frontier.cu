// dmv_frontier.cu -- DMV chart DP, "frontier" schedule: one CTA per sentence, one thread per (target cell, new terms).
//
// The DMV chart operator (reference /root/reference/src/model/torch_struct/dmv.py:19-66 and the autograd
// marginals / argmax of helpers.py:118-154).  The textbook evaluation of a width-w item is "gather w terms, reduce across
// lanes": shuffle trees, barriers and hundreds of instructions per warp on the critical path of every width, 2 (len)
// widths deep (round 1's role-split kernel, removed; dmv_gather.cu is the lean version of that idea).
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
// Shared memory per cell (diagonal-major index).  Every item kind is its OWN float2 array
// (.x = HASCHILD, .y = NOCHILD): neighbouring lanes own neighbouring cells of one width, so a warp's access to one array
// is 256 contiguous bytes -- two wavefronts, no bank conflicts -- and a task loads only the item kinds it needs.  (With
// one float4 per cell, 16-byte-strided scalar accesses were 4-way conflicted and half of every float4 load was unused:
// the reverse sweep was bound by shared-memory wavefronts, 57k of 111k clk for 40 words.)
//   log pass 88 B: CL, CR, IL, IR (IL / IR pre-loaded with attach + dec[GO]), X = (XL, XR), and 48 B that hold the running
//   (m, s) pairs while the state lives in shared memory, then the gradients gCL, gCR, gIL, gIR in the reverse sweep.
//   max pass 62 B: CL, CR, IL, IR, best-so-far VX = (XL, XR), VC = (CL.HAS, CL.NO, CR.HAS, CR.NO), 6 arg-max bytes.
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
// first cell of diagonal d: d Nb - d (d - 1) / 2 = d (2 Nb + 1 - d) / 2  (the product is always even; 2 Nb + 1 is loop-invariant)
__device__ __forceinline__ int dbase(int d, int Nb) { return (d * (2 * Nb + 1 - d)) >> 1; }
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
    float2 *CL, *CR, *IL, *IR, *X;   // values; X = (XL, XR) before the arc score is added
    float4 *A0, *A1, *A2;            // running (m, s) pairs of (XL, XR), (CL.HAS, CL.NO), (CR.HAS, CR.NO)
    float2 *gCL, *gCR, *gIL, *gIR;   // reverse sweep: gradients (alias A1 / A2)
};
struct MaxChart {
    float2 *CL, *CR, *IL, *IR, *VX;
    float4 *VC;
    uint8_t *bp;  // 6 bytes per cell: XL, XR, CL[HAS], CL[NO], CR[HAS], CR[NO]  (first maximal split)
};

// NT == 32 is the warp-per-sentence mode: the "block" of a sentence is one warp, its barrier is __syncwarp and its
// thread index the lane (several sentences share a CTA, each with its own slice of shared memory)
template <int NT>
__device__ __forceinline__ void blk_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}
template <int NT>
__device__ __forceinline__ int blk_tid() { return NT == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }

__device__ __forceinline__ int clamp_len(const DmvArgs &p, int b) {
    int len = b < p.n_len_inline ? (int)p.len_inline[b] : (int)p.lengths[b];
    return len < 0 ? 0 : (len > p.N - 1 ? p.N - 1 : len);
}

// stage dec, width-0 complete items (STOP decisions, dmv.py:39-40) and the arc scores attach + dec[GO]
// (formed first in fp32, exactly as dmv.py:36-37 does).  dec index = dir*4 + val*2 + decision.
// MODE 0: plain; 1: also publish the staged values in p.share (log CTA); 2: take them from p.share (max CTA)
// order-preserving float <-> int key (atomicMax on shared-memory ints gives the maximum of the floats)
__device__ __forceinline__ int fkey(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float funkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// mukey (log pass only, else null): per child position, the key of its best incoming arc score (first guess of the
// per-word offsets, see log_pass)
template <int NT, int MODE>
__device__ __forceinline__ void stage_inputs(const DmvArgs &p, int b, int Nb, float *sdec, uint16_t *cw, float2 *CL, float2 *CR,
                                             float2 *IL, float2 *IR, int *mukey = nullptr) {
    const int tid = blk_tid<NT>(), N = p.N;
    const float *dec = p.dec + (size_t)b * N * 8;
    const float *attach = p.attach + (size_t)b * N * N * 2;
    float *sh_dec = nullptr;
    float2 *sh_il = nullptr, *sh_ir = nullptr;
    if (MODE != 0) {
        sh_dec = p.share + (size_t)b * p.share_stride;
        sh_il = reinterpret_cast<float2 *>(sh_dec + N * 8);
        sh_ir = sh_il + ncells(N);
    }
    if (MODE == 2) {
        if (tid == 0) {
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.share_flag + b) : "memory");
                if (v != p.share_epoch) __nanosleep(100);
            } while (v != p.share_epoch);
        }
        blk_sync<NT>();
        dec = sh_dec;
    }
    // the first arc scores are requested together with dec, not after it: from host memory each is a PCIe round trip
    constexpr int kPre = 4;
    float2 pre[kPre];
    if (MODE != 2) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int t = tid + k * NT;
            pre[k] = make_float2(0.f, 0.f);
            if (t < Nb * Nb) {
                const int h = t / Nb, ch = t - h * Nb;
                pre[k] = *reinterpret_cast<const float2 *>(attach + ((size_t)h * N + ch) * 2);
            }
        }
    }
#pragma unroll 1
    for (int t = tid; t < Nb * 8; t += NT) {
        const float v = MODE == 2 ? __ldcg(dec + t) : dec[t];
        sdec[t] = v;
        if (MODE == 1) sh_dec[t] = v;
    }
    if (MODE != 2 && mukey)
        for (int k = tid; k < Nb; k += NT) mukey[k] = fkey(NEG_BIG);
#pragma unroll 1
    for (int d = tid; d < Nb; d += NT) {  // cell -> (width, left end)
        const int base = dbase(d, Nb);
        for (int lo = 0; lo < Nb - d; ++lo) cw[base + lo] = (uint16_t)((d << 8) | lo);
    }
    blk_sync<NT>();
#pragma unroll 1
    for (int i = tid; i < Nb; i += NT) {
        CL[i] = make_float2(sdec[i * 8 + 1], sdec[i * 8 + 3]);
        CR[i] = make_float2(sdec[i * 8 + 5], sdec[i * 8 + 7]);
    }
    if constexpr (MODE == 2) {
        const int nc = ncells(Nb);
#pragma unroll 1
        for (int c = Nb + tid; c < nc; c += NT) { IL[c] = __ldcg(sh_il + c); IR[c] = __ldcg(sh_ir + c); }
    } else {
    // arc scores in the order they lie in memory (row h: Nb contiguous float2), so that the reads coalesce -- they may
    // come straight from pinned host memory over PCIe (vlgae_dmv_parse_host)
#pragma unroll 1
    for (int t = tid, k = 0; t < Nb * Nb; t += NT, ++k) {
        const int h = t / Nb, ch = t - h * Nb;
        if (h == ch) continue;
        float2 a;
        if (k < kPre) a = k == 0 ? pre[0] : (k == 1 ? pre[1] : (k == 2 ? pre[2] : pre[3]));
        else a = *reinterpret_cast<const float2 *>(attach + ((size_t)h * N + ch) * 2);
        if (ch < h) {
            const float2 v = make_float2(__fadd_rn(a.x, sdec[h * 8 + 0]), __fadd_rn(a.y, sdec[h * 8 + 2]));
            const int c = cidx(ch, h - ch, Nb);
            IL[c] = v;
            if (MODE == 1) sh_il[c] = v;
            if (mukey) atomicMax(mukey + ch, fkey(fmaxf(v.x, v.y)));
        } else {
            const float2 v = make_float2(__fadd_rn(a.x, sdec[h * 8 + 4]), __fadd_rn(a.y, sdec[h * 8 + 6]));
            const int c = cidx(h, ch - h, Nb);
            IR[c] = v;
            if (MODE == 1) sh_ir[c] = v;
            if (mukey) atomicMax(mukey + ch, fkey(fmaxf(v.x, v.y)));
        }
    }
    if (MODE == 1) {
        __threadfence();
        blk_sync<NT>();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.share_flag + b), "r"(p.share_epoch) : "memory");
    }
    }
}

// ---------------------------------------------------------------------------------------------
// register-resident variants: thread t < H owns the target cells Nb + t + k H (k < CPT, H = cells / CPT rounded up) for
// the whole sweep and keeps their running state in registers -- no cell decode, no accumulator load / store, no task
// loop per phase.  The stride is H, not the block size: cells are ordered by width and a phase's band of widths
// (s .. 2s+1) never holds more than ~Nb^2 / 6 < H cells (CPT = 2), so a thread's two cells are never active in the same
// phase and the two visits never serialise on the phase's critical path.
// ---------------------------------------------------------------------------------------------
template <int NT, int CPT>
__device__ __forceinline__ void inside_reg(const LogChart &c, const uint16_t *cw, int Nb, int len, float mask_zero) {
    const int tid = blk_tid<NT>(), nc = ncells(Nb);
    const int H = (nc - Nb + CPT - 1) / CPT;
    int ow[CPT], oi[CPT];
    float ax[CPT][4], al[CPT][4], ar[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cc = Nb + tid + k * H;
        const int e = (tid < H && cc < nc) ? (int)cw[cc] : 0;
        ow[k] = e >> 8; oi[k] = e & 255;  // width 0 = no cell: never inside a band
#pragma unroll
        for (int q = 0; q < 4; ++q) { ax[k][q] = (q & 1) ? 0.f : NEG_BIG; al[k][q] = ax[k][q]; ar[k][q] = ax[k][q]; }
    }
    // ONE phase (one barrier) per width.  After phase s - 1 both the incomplete and the complete items of width s are final:
    // the complete item of a span needs, as its last-arriving term, the incomplete item of the SAME span (split r = i for
    // CL, r = j for CR), which the cell's owner has just produced itself -- so the owner folds that term in and finalises
    // CL / CR right behind IL / IR, without a barrier in between.  Phase s then folds I(s) into the complete targets of
    // width s + 1 .. 2 s - 1 and C(s) into the targets of width s + 1 .. 2 s + 1 (the two phases A(s), B(s) of the
    // two-barrier schedule, minus the same-span terms), and finalises width s + 1.
#pragma unroll 1
    for (int s = 0; s < len; ++s) {
        const int Ds = dbase(s, Nb);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int w = ow[k], i = oi[k];
            if ((unsigned)(w - 1 - s) <= (unsigned)s) {  // s + 1 <= w <= 2 s + 1
                const int j = i + w;
                if (w <= 2 * s - 1) {  // terms whose later operand is an incomplete item of width s (s + 1 <= w <= 2 s - 1)
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = c.CL[Dd + i].y;          // CL[i, j-s].NO
                    const float2 i3 = c.IL[Ds + j - s];        // IL[j-s, j]
                    const float2 i4 = c.IR[Ds + i];            // IR[i, i+s]
                    const float r4 = c.CR[Dd + i + s].y;       // CR[i+s, j].NO
                    lse1(al[k][0], al[k][1], l3 + i3.x);
                    lse1(al[k][2], al[k][3], l3 + i3.y);
                    lse1(ar[k][0], ar[k][1], i4.x + r4);
                    lse1(ar[k][2], ar[k][3], i4.y + r4);
                }
                const int De = dbase(w - 1 - s, Nb);
                // steps 1, 2 (dmv.py:50-56): XL (+)= CR[i,r].NO + CL[r+1,j].HAS, XR (+)= CR[i,r].HAS + CL[r+1,j].NO
                const float2 la = c.CR[Ds + i], ra = c.CL[De + i + s + 1];
                if (w - 1 - s != s) {
                    const float2 lb = c.CR[De + i], rb = c.CL[Ds + j - s];
                    lse2(ax[k][0], ax[k][1], la.y + ra.x, lb.y + rb.x);
                    lse2(ax[k][2], ax[k][3], la.x + ra.y, lb.x + rb.y);
                } else {
                    lse1(ax[k][0], ax[k][1], la.y + ra.x);
                    lse1(ax[k][2], ax[k][3], la.x + ra.y);
                }
                if (w <= 2 * s) {
                    const int Dd = De + Nb - (w - 1 - s);   // dbase(w - s)
                    const float l3 = c.CL[Ds + i].y;           // CL[i, i+s].NO
                    const float2 i3 = c.IL[Dd + i + s];        // IL[i+s, j]
                    const float2 i4 = c.IR[Dd + i];            // IR[i, j-s]
                    const float r4 = c.CR[Ds + j - s].y;       // CR[j-s, j].NO
                    lse1(al[k][0], al[k][1], l3 + i3.x);
                    lse1(al[k][2], al[k][3], l3 + i3.y);
                    lse1(ar[k][0], ar[k][1], i4.x + r4);
                    lse1(ar[k][2], ar[k][3], i4.y + r4);
                }
                if (w == s + 1) {  // this span's items are complete now: IL / IR, then CL / CR through the same-span terms
                    const int cc = Nb + tid + k * H;
                    const float xl = lse_fin(ax[k][0], ax[k][1]), xr = lse_fin(ax[k][2], ax[k][3]);
                    const float2 arcl = c.IL[cc], arcr = c.IR[cc];
                    const float2 il = make_float2(xl + arcl.x, xl + arcl.y), ir = make_float2(xr + arcr.x, xr + arcr.y);
                    const float l3 = c.CL[i].y, r4 = c.CR[j].y;  // CL[i, i].NO, CR[j, j].NO (width-0 cells)
                    lse1(al[k][0], al[k][1], l3 + il.x);
                    lse1(al[k][2], al[k][3], l3 + il.y);
                    lse1(ar[k][0], ar[k][1], ir.x + r4);
                    lse1(ar[k][2], ar[k][3], ir.y + r4);
                    float2 vr = make_float2(lse_fin(ar[k][0], ar[k][1]), lse_fin(ar[k][2], ar[k][3]));
                    if (i == 0 && w != len) vr = make_float2(mask_zero, mask_zero);  // single-root mask, dmv.py:63
                    c.IL[cc] = il; c.IR[cc] = ir;
                    c.X[cc] = make_float2(xl, xr);
                    c.CL[cc] = make_float2(lse_fin(al[k][0], al[k][1]), lse_fin(al[k][2], al[k][3]));
                    c.CR[cc] = vr;
                }
            }
        }
        blk_sync<NT>();
    }
}

// The same sweep in the LINEAR domain (sums of products on exp(offset scores), as csrc/dmv_gather.cu): the running state of
// a target is one float per item, a term one FMA -- no exp, no max, no log on the phase's chain, and a third of the
// instructions per visit, which is what a phase costs at 2-4 active warps per SM sub-partition.  The caller checks the
// range of the result and falls back to the log-domain sweep above.  X keeps 1 / X for the reverse sweep.
template <int NT, int CPT>
__device__ __forceinline__ void inside_lin_reg(const LogChart &c, const uint16_t *cw, int Nb, int len) {
    const int tid = blk_tid<NT>(), nc = ncells(Nb);
    const int H = (nc - Nb + CPT - 1) / CPT;
    int ow[CPT], oi[CPT];
    float2 ax[CPT], al[CPT], ar[CPT];  // (XL, XR), CL (HAS, NO), CR (HAS, NO)
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cc = Nb + tid + k * H;
        const int e = (tid < H && cc < nc) ? (int)cw[cc] : 0;
        ow[k] = e >> 8; oi[k] = e & 255;
        ax[k] = al[k] = ar[k] = make_float2(0.f, 0.f);
    }
#pragma unroll 1
    for (int s = 0; s < len; ++s) {
        const int Ds = dbase(s, Nb);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int w = ow[k], i = oi[k];
            if ((unsigned)(w - 1 - s) <= (unsigned)s) {  // s + 1 <= w <= 2 s + 1
                const int j = i + w;
                if (w <= 2 * s - 1) {
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = c.CL[Dd + i].y;
                    const float2 i3 = c.IL[Ds + j - s];
                    const float2 i4 = c.IR[Ds + i];
                    const float r4 = c.CR[Dd + i + s].y;
                    al[k].x = fmaf(l3, i3.x, al[k].x); al[k].y = fmaf(l3, i3.y, al[k].y);
                    ar[k].x = fmaf(i4.x, r4, ar[k].x); ar[k].y = fmaf(i4.y, r4, ar[k].y);
                }
                const int De = dbase(w - 1 - s, Nb);
                const float2 la = c.CR[Ds + i], ra = c.CL[De + i + s + 1];
                ax[k].x = fmaf(la.y, ra.x, ax[k].x);   // XL: CR.NO * CL.HAS
                ax[k].y = fmaf(la.x, ra.y, ax[k].y);   // XR: CR.HAS * CL.NO
                if (w - 1 - s != s) {
                    const float2 lb = c.CR[De + i], rb = c.CL[Ds + j - s];
                    ax[k].x = fmaf(lb.y, rb.x, ax[k].x);
                    ax[k].y = fmaf(lb.x, rb.y, ax[k].y);
                }
                if (w <= 2 * s) {
                    const int Dd = De + Nb - (w - 1 - s);
                    const float l3 = c.CL[Ds + i].y;
                    const float2 i3 = c.IL[Dd + i + s];
                    const float2 i4 = c.IR[Dd + i];
                    const float r4 = c.CR[Ds + j - s].y;
                    al[k].x = fmaf(l3, i3.x, al[k].x); al[k].y = fmaf(l3, i3.y, al[k].y);
                    ar[k].x = fmaf(i4.x, r4, ar[k].x); ar[k].y = fmaf(i4.y, r4, ar[k].y);
                }
                if (w == s + 1) {
                    const int cc = Nb + tid + k * H;
                    const float2 arcl = c.IL[cc], arcr = c.IR[cc];  // linear arc factors
                    const float2 il = make_float2(ax[k].x * arcl.x, ax[k].x * arcl.y), ir = make_float2(ax[k].y * arcr.x, ax[k].y * arcr.y);
                    const float l3 = c.CL[i].y, r4 = c.CR[j].y;
                    float2 vl = make_float2(fmaf(l3, il.x, al[k].x), fmaf(l3, il.y, al[k].y));
                    float2 vr = make_float2(fmaf(ir.x, r4, ar[k].x), fmaf(ir.y, r4, ar[k].y));
                    if (i == 0 && w != len) vr = make_float2(0.f, 0.f);  // single-root mask, dmv.py:63 (exp(-1e12))
                    c.IL[cc] = il; c.IR[cc] = ir;
                    c.X[cc] = make_float2(ax[k].x > 0.f ? __fdividef(1.f, ax[k].x) : 0.f, ax[k].y > 0.f ? __fdividef(1.f, ax[k].y) : 0.f);
                    c.CL[cc] = vl;
                    c.CR[cc] = vr;
                }
            }
        }
        blk_sync<NT>();
    }
}

template <int NT, int CPT>
__device__ __forceinline__ void viterbi_reg(const MaxChart &c, const uint16_t *cw, int Nb, int len, float mask_zero) {
    const int tid = blk_tid<NT>(), nc = ncells(Nb);
    const int H = (nc - Nb + CPT - 1) / CPT;
    int ow[CPT], oi[CPT];
    float vx[CPT][2], vc[CPT][4];
    int bx[CPT][2], bc[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int cc = Nb + tid + k * H;
        const int e = (tid < H && cc < nc) ? (int)cw[cc] : 0;
        ow[k] = e >> 8; oi[k] = e & 255;
        vx[k][0] = vx[k][1] = NEG_BIG; bx[k][0] = bx[k][1] = 255;
#pragma unroll
        for (int q = 0; q < 4; ++q) { vc[k][q] = NEG_BIG; bc[k][q] = 255; }
    }
    // one phase per width, as in inside_reg: the owner of a span finalises IL / IR and, through the same-span terms, CL / CR
    // in the same visit (first-maximum ties are settled by the split index, so the order of arrival does not matter)
#pragma unroll 1
    for (int s = 0; s < len; ++s) {
        const int Ds = dbase(s, Nb);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int w = ow[k], i = oi[k];
            if ((unsigned)(w - 1 - s) <= (unsigned)s) {  // s + 1 <= w <= 2 s + 1
                const int j = i + w;
                if (w <= 2 * s - 1) {  // later operand = an incomplete item of width s
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = c.CL[Dd + i].y;
                    const float2 i3 = c.IL[Ds + j - s];
                    const float2 i4 = c.IR[Ds + i];
                    const float r4 = c.CR[Dd + i + s].y;
                    amax1(vc[k][0], bc[k][0], __fadd_rn(l3, i3.x), w - s);   // CL split r - i, r = j - s
                    amax1(vc[k][1], bc[k][1], __fadd_rn(l3, i3.y), w - s);
                    amax1(vc[k][2], bc[k][2], __fadd_rn(i4.x, r4), s - 1);   // CR split r - i - 1, r = i + s
                    amax1(vc[k][3], bc[k][3], __fadd_rn(i4.y, r4), s - 1);
                }
                const int De = dbase(w - 1 - s, Nb);
                const float2 la = c.CR[Ds + i], ra = c.CL[De + i + s + 1];
                amax1(vx[k][0], bx[k][0], __fadd_rn(la.y, ra.x), s);
                amax1(vx[k][1], bx[k][1], __fadd_rn(la.x, ra.y), s);
                if (w - 1 - s != s) {
                    const float2 lb = c.CR[De + i], rb = c.CL[Ds + j - s];
                    amax1(vx[k][0], bx[k][0], __fadd_rn(lb.y, rb.x), w - 1 - s);
                    amax1(vx[k][1], bx[k][1], __fadd_rn(lb.x, rb.y), w - 1 - s);
                }
                if (w <= 2 * s) {
                    const int Dd = De + Nb - (w - 1 - s);   // dbase(w - s)
                    const float l3 = c.CL[Ds + i].y;
                    const float2 i3 = c.IL[Dd + i + s];
                    const float2 i4 = c.IR[Dd + i];
                    const float r4 = c.CR[Ds + j - s].y;
                    amax1(vc[k][0], bc[k][0], __fadd_rn(l3, i3.x), s);
                    amax1(vc[k][1], bc[k][1], __fadd_rn(l3, i3.y), s);
                    amax1(vc[k][2], bc[k][2], __fadd_rn(i4.x, r4), w - s - 1);
                    amax1(vc[k][3], bc[k][3], __fadd_rn(i4.y, r4), w - s - 1);
                }
                if (w == s + 1) {
                    const int cc = Nb + tid + k * H;
                    const float2 arcl = c.IL[cc], arcr = c.IR[cc];
                    const float2 il = make_float2(__fadd_rn(vx[k][0], arcl.x), __fadd_rn(vx[k][0], arcl.y));
                    const float2 ir = make_float2(__fadd_rn(vx[k][1], arcr.x), __fadd_rn(vx[k][1], arcr.y));
                    const float l3 = c.CL[i].y, r4 = c.CR[j].y;  // CL[i, i].NO, CR[j, j].NO
                    amax1(vc[k][0], bc[k][0], __fadd_rn(l3, il.x), 0);       // CL split 0 (r = i)
                    amax1(vc[k][1], bc[k][1], __fadd_rn(l3, il.y), 0);
                    amax1(vc[k][2], bc[k][2], __fadd_rn(ir.x, r4), w - 1);   // CR split w - 1 (r = j)
                    amax1(vc[k][3], bc[k][3], __fadd_rn(ir.y, r4), w - 1);
                    float2 vr = make_float2(vc[k][2], vc[k][3]);
                    if (i == 0 && w != len) vr = make_float2(mask_zero, mask_zero);
                    c.IL[cc] = il; c.IR[cc] = ir;
                    c.CL[cc] = make_float2(vc[k][0], vc[k][1]);
                    c.CR[cc] = vr;
                    uint8_t *bp = c.bp + cc * 6;
                    bp[0] = (uint8_t)bx[k][0]; bp[1] = (uint8_t)bx[k][1];
                    bp[2] = (uint8_t)bc[k][0]; bp[3] = (uint8_t)bc[k][1]; bp[4] = (uint8_t)bc[k][2]; bp[5] = (uint8_t)bc[k][3];
                }
            }
        }
        blk_sync<NT>();
    }
}

// ---------------------------------------------------------------------------------------------
// log semiring: inside + outside for one sentence
// ---------------------------------------------------------------------------------------------
// `small` = shared memory for dec and the cell table; `chart` = the chart arrays, in shared memory right behind them
// (null) or, for sentences whose chart does not fit, in this CTA's slice of the global workspace (L2-resident).
// (GC is a template parameter so that the shared-memory case keeps its address space at compile time: through one generic
// pointer every chart access became a generic load / store and the cfg2 launch went from 56 to 60 us)
// dec | cell table | per-word offsets (Nb + 1 floats, log pass)
__host__ __device__ inline size_t small_bytes(int Nb) {
    return (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15) + (((size_t)ncells(Nb) * 2 + 15) & ~(size_t)15) +
           (((size_t)(Nb + 1) * 4 + 15) & ~(size_t)15);
}
template <bool GC>
__device__ __forceinline__ unsigned char *chart_base(unsigned char *small, unsigned char *chart, int Nb) {
    if (GC) return chart;
    return small + small_bytes(Nb);
}

// LIN: the linear-domain variant (register-state variants with the chart in shared memory only).  It returns false -- having
// written nothing that the log-domain pass does not overwrite -- when the result left the safe fp32 range or failed its
// self-check; the caller then runs the log-domain pass.
template <int NT, int CPT, bool GC, bool LIN>
__device__ bool log_pass_impl(const DmvArgs &p, int b, int len, unsigned char *small, unsigned char *chart) {
    const int tid = blk_tid<NT>(), N = p.N;
    const int Nb = len + 1, nc = ncells(Nb);
    float *sdec = reinterpret_cast<float *>(small);
    uint16_t *cw = reinterpret_cast<uint16_t *>(small + (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15));
    LogChart c;
    unsigned char *base = chart_base<GC>(small, chart, Nb);
    if (CPT > 0) {  // running state in registers: only the 32 B of gradients per cell
        c.A0 = c.A1 = c.A2 = nullptr;
        c.gCL = reinterpret_cast<float2 *>(base); c.gCR = c.gCL + nc; c.gIL = c.gCR + nc; c.gIR = c.gIL + nc;
        c.CL = c.gIR + nc;
    } else {
        c.A0 = reinterpret_cast<float4 *>(base);
        c.A1 = c.A0 + nc; c.A2 = c.A1 + nc;
        c.gCL = reinterpret_cast<float2 *>(c.A1); c.gCR = c.gCL + nc;
        c.gIL = reinterpret_cast<float2 *>(c.A2); c.gIR = c.gIL + nc;
        c.CL = reinterpret_cast<float2 *>(c.A2 + nc);
    }
    c.CR = c.CL + nc; c.IL = c.CR + nc; c.IR = c.IL + nc; c.X = c.IR + nc;
    const bool want_grad = (p.gdec != nullptr) || (p.gattach != nullptr);
    const bool prof = p.prof && b == 0 && tid == 0;
    long long t0c = 0;
    if (prof) t0c = clock64();
    constexpr bool reg_state = CPT > 0;  // the launcher picks CPT so that the sentence's cells fit (cap - 1 words)

    float *mu = reinterpret_cast<float *>(small + (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15) + (((size_t)nc * 2 + 15) & ~(size_t)15));
    float zres = 0.f, ztop = 1.f;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
    int *mukey = attempt == 0 ? reinterpret_cast<int *>(mu) : nullptr;
    if (p.share && attempt == 0) stage_inputs<NT, 1>(p, b, Nb, sdec, cw, c.CL, c.CR, c.IL, c.IR, mukey);
    else stage_inputs<NT, 0>(p, b, Nb, sdec, cw, c.CL, c.CR, c.IL, c.IR, mukey);
    blk_sync<NT>();
    // Per-word offsets: every arc score into word k is lowered by mu[k]; every tree then scores sum_k mu[k] less, so the
    // posteriors are unchanged and log Z is restored at the end, while the chart values stay O(10) instead of
    // O(-4 len) -- one fp32 ulp there is 1e-6 instead of 1.5e-5, and the marginals come out ~10x closer to the exact
    // ones than the reference's own fp32 sweep.  First guess: best incoming arc (collected while staging) + mean STOP
    // costs of the word; if the top of the chart still ends up far from 0, the sweep is repeated once with the guess
    // corrected by its residual.
    if (tid < 32) {
        float part = 0.f;
        for (int k = tid; k < Nb; k += 32) {
            float u = 0.f;
            if (attempt == 0) {
                const float m = funkey(mukey[k]);
                if (k >= 1 && m > -1e6f && !p.no_offsets)
                    u = m + 0.5f * (sdec[k * 8 + 1] + sdec[k * 8 + 3]) + 0.5f * (sdec[k * 8 + 5] + sdec[k * 8 + 7]);
                if (!(fabsf(u) < 1e6f)) u = 0.f;
            } else if (k >= 1) {
                u = mu[k] + zres / (float)len;
            }
            mu[k] = u;
            part += u;
        }
        for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (tid == 0) mu[Nb] = part;
    }
    blk_sync<NT>();
#pragma unroll 1
    for (int cc = Nb + tid; cc < nc; cc += NT) {  // cell (width d, left end lo): IL is the arc into lo, IR the arc into lo + d
        const int d = cw[cc] >> 8, lo = cw[cc] & 255;
        const float ml = mu[lo], mr = mu[lo + d];
        const float2 vl = c.IL[cc], vr = c.IR[cc];
        if (LIN) {  // linear arc factors (a score of -1e12 / -1e20 becomes an exact 0)
            c.IL[cc] = make_float2(fexp(__fadd_rn(vl.x, -ml)), fexp(__fadd_rn(vl.y, -ml)));
            c.IR[cc] = make_float2(fexp(__fadd_rn(vr.x, -mr)), fexp(__fadd_rn(vr.y, -mr)));
        } else {
            c.IL[cc] = make_float2(__fadd_rn(vl.x, -ml), __fadd_rn(vl.y, -ml));
            c.IR[cc] = make_float2(__fadd_rn(vr.x, -mr), __fadd_rn(vr.y, -mr));
        }
    }
    if (LIN) {
#pragma unroll 1
        for (int i = tid; i < Nb; i += NT) {  // STOP factors of the width-0 items
            const float2 l = c.CL[i], r = c.CR[i];
            c.CL[i] = make_float2(fexp(l.x), fexp(l.y));
            c.CR[i] = make_float2(fexp(r.x), fexp(r.y));
        }
    }
    if (!reg_state) {
        const float4 init = make_float4(NEG_BIG, 0.f, NEG_BIG, 0.f);
#pragma unroll 1
        for (int t = Nb + tid; t < nc; t += NT) { c.A0[t] = init; c.A1[t] = init; c.A2[t] = init; }
    }
    blk_sync<NT>();
    if (prof) p.prof[0] = clock64() - t0c;

    // ---------------- inside ----------------
    if (LIN) {
        inside_lin_reg<NT, (CPT > 0 ? CPT : 1)>(c, cw, Nb, len);
    } else if (reg_state) {
        inside_reg<NT, (CPT > 0 ? CPT : 1)>(c, cw, Nb, len, p.mask_zero);
    } else {
        // one phase per width, as in inside_reg (the span's owner finalises IL / IR and, through the same-span terms,
        // CL / CR in one visit): a visit loads and stores the running state of its cell once for all three kinds of terms
#pragma unroll 1
        for (int s = 0; s < len; ++s) {
            const int Ds = dbase(s, Nb);
            const int c0 = dbase(s + 1, Nb), c1 = dbase(min(2 * s + 1, len) + 1, Nb);  // targets of width s + 1 .. 2 s + 1
#pragma unroll 1
            for (int cc = c0 + tid; cc < c1; cc += NT) {
                const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                const int De = dbase(w - 1 - s, Nb);
                float4 a0 = c.A0[cc], a1, a2;
                const bool ctarget = w <= 2 * s || w == s + 1;  // complete-item terms arrive, or the span is finalised
                if (ctarget) { a1 = c.A1[cc]; a2 = c.A2[cc]; }
                if (w <= 2 * s - 1) {  // later operand = an incomplete item of width s
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = c.CL[Dd + i].y;          // CL[i, j-s].NO
                    const float2 i3 = c.IL[Ds + j - s];        // IL[j-s, j]
                    const float2 i4 = c.IR[Ds + i];            // IR[i, i+s]
                    const float r4 = c.CR[Dd + i + s].y;       // CR[i+s, j].NO
                    lse1(a1.x, a1.y, l3 + i3.x);
                    lse1(a1.z, a1.w, l3 + i3.y);
                    lse1(a2.x, a2.y, i4.x + r4);
                    lse1(a2.z, a2.w, i4.y + r4);
                }
                // steps 1, 2 (dmv.py:50-56): XL (+)= CR[i,r].NO + CL[r+1,j].HAS, XR (+)= CR[i,r].HAS + CL[r+1,j].NO
                const float2 la = c.CR[Ds + i], ra = c.CL[De + i + s + 1];
                if (w - 1 - s != s) {
                    const float2 lb = c.CR[De + i], rb = c.CL[Ds + j - s];
                    lse2(a0.x, a0.y, la.y + ra.x, lb.y + rb.x);
                    lse2(a0.z, a0.w, la.x + ra.y, lb.x + rb.y);
                } else {
                    lse1(a0.x, a0.y, la.y + ra.x);
                    lse1(a0.z, a0.w, la.x + ra.y);
                }
                if (w <= 2 * s) {
                    const int Dd = De + Nb - (w - 1 - s);   // dbase(w - s)
                    const float l3 = c.CL[Ds + i].y;        // CL[i, i+s].NO
                    const float2 i3 = c.IL[Dd + i + s];     // IL[i+s, j]
                    const float2 i4 = c.IR[Dd + i];         // IR[i, j-s]
                    const float r4 = c.CR[Ds + j - s].y;    // CR[j-s, j].NO
                    lse1(a1.x, a1.y, l3 + i3.x);
                    lse1(a1.z, a1.w, l3 + i3.y);
                    lse1(a2.x, a2.y, i4.x + r4);
                    lse1(a2.z, a2.w, i4.y + r4);
                }
                if (w == s + 1) {
                    const float xl = lse_fin(a0.x, a0.y), xr = lse_fin(a0.z, a0.w);
                    const float2 arcl = c.IL[cc], arcr = c.IR[cc];
                    const float2 il = make_float2(xl + arcl.x, xl + arcl.y), ir = make_float2(xr + arcr.x, xr + arcr.y);
                    const float l3 = c.CL[i].y, r4 = c.CR[j].y;  // CL[i, i].NO, CR[j, j].NO (width-0 cells)
                    lse1(a1.x, a1.y, l3 + il.x);
                    lse1(a1.z, a1.w, l3 + il.y);
                    lse1(a2.x, a2.y, ir.x + r4);
                    lse1(a2.z, a2.w, ir.y + r4);
                    float2 vr = make_float2(lse_fin(a2.x, a2.y), lse_fin(a2.z, a2.w));
                    if (i == 0 && w != len) vr = make_float2(p.mask_zero, p.mask_zero);  // single-root mask, dmv.py:63
                    c.IL[cc] = il; c.IR[cc] = ir;
                    c.X[cc] = make_float2(xl, xr);
                    c.CL[cc] = make_float2(lse_fin(a1.x, a1.y), lse_fin(a1.z, a1.w));
                    c.CR[cc] = vr;
                } else {
                    c.A0[cc] = a0;
                    if (ctarget) { c.A1[cc] = a1; c.A2[cc] = a2; }
                }
            }
            blk_sync<NT>();
        }
    }
    if (LIN) {
        ztop = c.CR[cidx(0, len, Nb)].y;          // linear domain
        if (!(ztop > 1e-30f && ztop < 1e30f)) {   // out of the safe range (also NaN): the log-domain pass takes over
            blk_sync<NT>();
            return false;
        }
        zres = __logf(ztop);
        // (exp(chart value) stays far inside the fp32 range for |log Z'| of this size; n = 64 x 512: 654 k -> 804 k sentences/s
        // against a fixed 16, same 2.3e-7 from fp64)
        if (fabsf(zres) <= (p.retry_above > 0.f ? p.retry_above : fmaxf(16.f, 0.5f * (float)len)) || len == 0) break;
        blk_sync<NT>();
        continue;
    }
    zres = c.CR[cidx(0, len, Nb)].y;  // dmv.py:65, minus the offsets
    // (chart values of magnitude <= max(48, len) keep one fp32 ulp <= 8e-6; n = 128 without the second sweep: |gpu - fp64|
    // 5.5e-6 instead of 1.8e-6, 6.9 instead of 8.9 ms at B = 512)
    if (fabsf(zres) <= (p.retry_above > 0.f ? p.retry_above : fmaxf(48.f, (float)len)) || len == 0) break;
    blk_sync<NT>();
    }
    if (prof) p.prof[1] = clock64() - t0c;
    if (LIN && !(fabsf(zres) <= 40.f)) { blk_sync<NT>(); return false; }  // the corrected offsets did not bring Z' near 1
    if (!LIN || !want_grad) {
        if (tid == 0) p.Z[b] = zres + mu[Nb];
    }
    if (!want_grad) { blk_sync<NT>(); return true; }

    // ---------------- outside (explicit reverse sweep) ----------------
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *g4 = reinterpret_cast<float4 *>(c.gCL);  // gCL, gCR, gIL, gIR are contiguous: 2 nc float4
#pragma unroll 1
        for (int t = tid; t < 2 * nc; t += NT) g4[t] = z;
        blk_sync<NT>();
        if (tid == 0) c.gCR[cidx(0, len, Nb)].y = LIN ? __fdividef(p.gZ ? p.gZ[b] : 1.f, ztop) : (p.gZ ? p.gZ[b] : 1.f);
        blk_sync<NT>();
    }
    unsigned rw_next = __float2uint_rz(__fdividef(1048576.0f, (float)(Nb - len))) + 2u;
#pragma unroll 1
    for (int w = len; w >= 1; --w) {
        const int ntask = (Nb - w) * w;
        // task t = (split a, parent i) with i fastest: neighbouring lanes touch neighbouring cells of one width in
        // every array.  All accumulator words of a task are distinct and are loaded BEFORE the first store (with
        // interleaved read-modify-writes every load has to stay behind the previous store: possible aliasing).
        const int np = Nb - w;
        // t / np == (t * rw) >> 20 needs rw * np >= 2^20 and t * (rw * np - 2^20) < 2^20; the float quotient is within 1
        // of floor(2^20 / np), so +2 keeps the first and 3 np * t <= 3 * 75 * 1406 the second (no integer division)
        // (computed one width ahead: the conversion + reciprocal chain stays off the head of the phase)
        const unsigned rw = rw_next;
        rw_next = __float2uint_rz(__fdividef(1048576.0f, (float)(np + 1))) + 2u;
        const int pb = dbase(w, Nb);
        if (LIN) {
            // linear adjoints (alpha_P += alpha_L alpha_R  =>  beta_L += beta_P alpha_R, beta_R += beta_P alpha_L): the same
            // tasks and the same single writer per accumulator word, one FMA per term instead of an exponential
#pragma unroll 1
            for (int t = tid; t < ntask; t += NT) {
                const int a = (int)(((unsigned)t * rw) >> 20), i = t - a * np;
                const float2 bl = c.gCL[pb + i];
                float2 br = c.gCR[pb + i];
                if (i == 0 && w != len) br = make_float2(0.f, 0.f);  // the mask overwrote CR[0][w]: nothing passes
                const int cl3 = cidx(i, a, Nb), cr3 = cidx(i + a, w - a, Nb);
                const int cl4 = cl3 + Nb - a, cr4 = cr3 - Nb + w - a;
                const float lv3 = c.CL[cl3].y;
                const float2 rv3 = c.IL[cr3];
                const float2 lv4 = c.IR[cl4];
                const float rv4 = c.CR[cr4].y;
                const float o1 = c.gCL[cl3].y, o4 = c.gCR[cr4].y;
                const float2 o2 = c.gIL[cr3], o3 = c.gIR[cl4];
                c.gCL[cl3].y = fmaf(bl.x, rv3.x, fmaf(bl.y, rv3.y, o1));
                c.gIL[cr3] = make_float2(fmaf(bl.x, lv3, o2.x), fmaf(bl.y, lv3, o2.y));
                c.gIR[cl4] = make_float2(fmaf(br.x, rv4, o3.x), fmaf(br.y, rv4, o3.y));
                c.gCR[cr4].y = fmaf(br.x, lv4.x, fmaf(br.y, lv4.y, o4));
            }
            blk_sync<NT>();
#pragma unroll 1
            for (int t = tid; t < ntask; t += NT) {
                const int a = (int)(((unsigned)t * rw) >> 20), i = t - a * np;
                const float2 gil = c.gIL[pb + i], gir = c.gIR[pb + i], il = c.IL[pb + i], ir = c.IR[pb + i], ix = c.X[pb + i];
                // beta X = sum_v beta I[v] * arc[v],  arc[v] = I[v] / X
                const float bxl = fmaf(gil.x, il.x, gil.y * il.y) * ix.x, bxr = fmaf(gir.x, ir.x, gir.y * ir.y) * ix.y;
                const int cl = cidx(i, a, Nb), cr = cidx(i + a + 1, w - 1 - a, Nb);
                const float2 lv = c.CR[cl], rv = c.CL[cr];
                const float2 ol = c.gCR[cl], orr = c.gCL[cr];
                c.gCR[cl] = make_float2(fmaf(bxr, rv.y, ol.x), fmaf(bxl, rv.x, ol.y));    // CR.HAS from XR (x CL.NO), CR.NO from XL (x CL.HAS)
                c.gCL[cr] = make_float2(fmaf(bxl, lv.y, orr.x), fmaf(bxr, lv.x, orr.y));  // CL.HAS from XL (x CR.NO), CL.NO from XR (x CR.HAS)
            }
            blk_sync<NT>();
            continue;
        }
        // phase A'(w): complete parents of width w (steps 3, 4 transposed)
#pragma unroll 1
        for (int t = tid; t < ntask; t += NT) {
            const int a = (int)(((unsigned)t * rw) >> 20), i = t - a * np;
            const float2 gl = c.gCL[pb + i], gr = c.gCR[pb + i], outl = c.CL[pb + i], outr = c.CR[pb + i];
            const int cl3 = cidx(i, a, Nb), cr3 = cidx(i + a, w - a, Nb);
            // dbase(d + 1) - dbase(d) = Nb - d:  cl4 = cidx(i, a + 1), cr4 = cidx(i + a + 1, w - 1 - a)
            const int cl4 = cl3 + Nb - a, cr4 = cr3 - Nb + w - a;
            const float lv3 = c.CL[cl3].y;
            const float2 rv3 = c.IL[cr3];
            const float2 lv4 = c.IR[cl4];
            const float rv4 = c.CR[cr4].y;
            const float o1 = c.gCL[cl3].y, o4 = c.gCR[cr4].y;
            const float2 o2 = c.gIL[cr3], o3 = c.gIR[cl4];
            const float p0 = gl.x * fexp(lv3 + rv3.x - outl.x), p1 = gl.y * fexp(lv3 + rv3.y - outl.y);
            float q0 = 0.f, q1 = 0.f;
            if (!(i == 0 && w != len)) {  // the mask overwrote CR[0][w]: no gradient passes through it
                q0 = gr.x * fexp(lv4.x + rv4 - outr.x);
                q1 = gr.y * fexp(lv4.y + rv4 - outr.y);
            }
            c.gCL[cl3].y = o1 + (p0 + p1);
            c.gIL[cr3] = make_float2(o2.x + p0, o2.y + p1);
            c.gIR[cl4] = make_float2(o3.x + q0, o3.y + q1);
            c.gCR[cr4].y = o4 + (q0 + q1);
        }
        blk_sync<NT>();
        // phase B'(w): incomplete parents of width w (steps 1, 2 transposed)
#pragma unroll 1
        for (int t = tid; t < ntask; t += NT) {
            const int a = (int)(((unsigned)t * rw) >> 20), i = t - a * np;
            const float2 gil = c.gIL[pb + i], gir = c.gIR[pb + i], x = c.X[pb + i];
            const float gl = gil.x + gil.y, gr = gir.x + gir.y;
            const int cl = cidx(i, a, Nb), cr = cidx(i + a + 1, w - 1 - a, Nb);
            const float2 lv = c.CR[cl], rv = c.CL[cr];
            const float2 ol = c.gCR[cl], orr = c.gCL[cr];
            const float pL = gl * fexp(lv.y + rv.x - x.x), pR = gr * fexp(lv.x + rv.y - x.y);
            c.gCR[cl] = make_float2(ol.x + pR, ol.y + pL);    // CR[i, r]: .HAS from step 2, .NO from step 1
            c.gCL[cr] = make_float2(orr.x + pL, orr.y + pR);  // CL[r+1, j]: .HAS from step 1, .NO from step 2
        }
        blk_sync<NT>();
    }
    if (prof) p.prof[2] = clock64() - t0c;

    // ---------------- outputs ----------------
    if (LIN) {
        // alpha * beta = d (gZ log Z) / d (log potential), in place; then the self-check: every word has exactly one head, so
        // its arc marginals sum to gZ
#pragma unroll 1
        for (int cc = tid; cc < nc; cc += NT) {
            if (cc < Nb) {
                const float2 a = c.CL[cc], g = c.gCL[cc], a2 = c.CR[cc], g2 = c.gCR[cc];
                c.gCL[cc] = make_float2(a.x == 0.f ? 0.f : a.x * g.x, a.y == 0.f ? 0.f : a.y * g.y);
                c.gCR[cc] = make_float2(a2.x == 0.f ? 0.f : a2.x * g2.x, a2.y == 0.f ? 0.f : a2.y * g2.y);
            } else {
                const float2 a = c.IL[cc], g = c.gIL[cc], a2 = c.IR[cc], g2 = c.gIR[cc];
                c.gIL[cc] = make_float2(a.x == 0.f ? 0.f : a.x * g.x, a.y == 0.f ? 0.f : a.y * g.y);
                c.gIR[cc] = make_float2(a2.x == 0.f ? 0.f : a2.x * g2.x, a2.y == 0.f ? 0.f : a2.y * g2.y);
            }
        }
        blk_sync<NT>();
        const float gz = p.gZ ? p.gZ[b] : 1.f;
        bool good = true;
#pragma unroll 1
        for (int k = 1 + tid; k < Nb; k += NT) {
            float t0 = 0.f;
            for (int h = 0; h < k; ++h) { const float2 v = c.gIR[cidx(h, k - h, Nb)]; t0 += v.x + v.y; }       // heads to the left
            for (int h = k + 1; h < Nb; ++h) { const float2 v = c.gIL[cidx(k, h - k, Nb)]; t0 += v.x + v.y; }  // heads to the right
            if (!(fabsf(t0 - gz) <= 1e-3f * fabsf(gz) + 1e-30f)) good = false;
        }
        if (NT == 32) good = __all_sync(0xffffffffu, good);
        else good = __syncthreads_and(good);
        if (!good) return false;
        if (tid == 0) p.Z[b] = zres + mu[Nb];
    }
    if (p.gattach) {
        float2 *ga = reinterpret_cast<float2 *>(p.gattach + (size_t)b * N * N * 2);
#pragma unroll 1
        for (int t = tid; t < N * N; t += NT) {
            const int h = t / N, ch = t - h * N;
            float2 v = make_float2(0.f, 0.f);
            if (h < Nb && ch < Nb && h != ch) v = ch < h ? c.gIL[cidx(ch, h - ch, Nb)] : c.gIR[cidx(h, ch - h, Nb)];
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
                if (dir == 0) {
                    for (int ch = 0; ch < i; ++ch) { const float2 v = c.gIL[cidx(ch, i - ch, Nb)]; go.x += v.x; go.y += v.y; }
                    stop = c.gCL[i];
                } else {
                    for (int d = 1; d < Nb - i; ++d) { const float2 v = c.gIR[cidx(i, d, Nb)]; go.x += v.x; go.y += v.y; }
                    stop = c.gCR[i];
                }
            }
            *reinterpret_cast<float4 *>(gd + i * 8 + dir * 4) = make_float4(go.x, stop.x, go.y, stop.y);  // [dir][val][decision]
        }
    }
    blk_sync<NT>();
    if (prof) p.prof[3] = clock64() - t0c;
    return true;
}

template <int NT, int CPT, bool GC>
__device__ __forceinline__ void log_pass(const DmvArgs &p, int b, int len, unsigned char *small, unsigned char *chart) {
    if constexpr (CPT > 0 && !GC) {
        if ((len <= p.lin_max_len || (p.lin_long_from > 0 && len >= p.lin_long_from)) &&
            log_pass_impl<NT, CPT, GC, true>(p, b, len, small, chart)) return;
    }
    log_pass_impl<NT, CPT, GC, false>(p, b, len, small, chart);
}

// ---------------------------------------------------------------------------------------------
// max semiring: Viterbi chart with first-max back-pointers + breadth-first back-trace
// ---------------------------------------------------------------------------------------------
// items of the back-trace: kind (0 CR, 1 CL, 2 IR, 3 IL) | v << 2 | lo << 3 | hi << 12
__device__ __forceinline__ int mk_item(int kind, int v, int lo, int hi) { return kind | (v << 2) | (lo << 3) | (hi << 12); }

template <int NT, int CPT, bool GC>
__device__ void max_pass(const DmvArgs &p, int b, int len, unsigned char *small, unsigned char *chart) {
    const int tid = blk_tid<NT>(), N = p.N;
    const int Nb = len + 1, nc = ncells(Nb);
    float *sdec = reinterpret_cast<float *>(small);
    uint16_t *cw = reinterpret_cast<uint16_t *>(small + (((size_t)Nb * 8 * 4 + 15) & ~(size_t)15));
    MaxChart c;
    c.VC = reinterpret_cast<float4 *>(chart_base<GC>(small, chart, Nb));
    c.CL = reinterpret_cast<float2 *>(c.VC + nc);
    c.CR = c.CL + nc; c.IL = c.CR + nc; c.IR = c.IL + nc; c.VX = c.IR + nc;
    int *queue = reinterpret_cast<int *>(c.VX + nc);  // 2 x (2 Nb + 2) ints
    c.bp = reinterpret_cast<uint8_t *>(queue + 2 * (2 * Nb + 2));
    const bool prof = p.prof && b == 0 && tid == 0;
    long long t0c = 0;
    if (prof) t0c = clock64();
    constexpr bool reg_state = CPT > 0;

    if (p.share) stage_inputs<NT, 2>(p, b, Nb, sdec, cw, c.CL, c.CR, c.IL, c.IR);
    else stage_inputs<NT, 0>(p, b, Nb, sdec, cw, c.CL, c.CR, c.IL, c.IR);
    if (!reg_state) {
#pragma unroll 1
        for (int t = Nb + tid; t < nc; t += NT) {
            c.VC[t] = make_float4(NEG_BIG, NEG_BIG, NEG_BIG, NEG_BIG);
            c.VX[t] = make_float2(NEG_BIG, NEG_BIG);
        }
#pragma unroll 1
        for (int t = tid; t < nc * 6; t += NT) c.bp[t] = 255;
    }
    // outputs that the back-trace only dots with ones are zero-filled up front
    if (p.arcs) {
        float2 *z = reinterpret_cast<float2 *>(p.arcs + (size_t)b * N * N * 2);
#pragma unroll 1
        for (int t = tid; t < N * N; t += NT) z[t] = make_float2(0.f, 0.f);
    }
    if (p.vgdec) for (int t = tid; t < N * 8; t += NT) p.vgdec[(size_t)b * N * 8 + t] = 0.f;
    if (p.heads) for (int t = tid; t < N; t += NT) p.heads[(size_t)b * N + t] = 0;
    blk_sync<NT>();
    if (prof) p.prof[4] = clock64() - t0c;

    if (reg_state) {
        viterbi_reg<NT, (CPT > 0 ? CPT : 1)>(c, cw, Nb, len, p.mask_zero);
    } else {
#pragma unroll 1
        for (int s = 0; s <= len; ++s) {
            const int Ds = dbase(s, Nb);
            if (s >= 1) {
                // phase A(s): the new term of CL is split r - i = w - s, of CR split r - i - 1 = s - 1
                const int whi = min(2 * s - 1, len);
                const int c1 = dbase(whi + 1, Nb);
#pragma unroll 1
                for (int cc = Ds + tid; cc < c1; cc += NT) {
                    const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                    const int Dd = dbase(w - s, Nb);
                    const float l3 = c.CL[Dd + i].y;
                    const float2 i3 = c.IL[Ds + j - s];
                    const float2 i4 = c.IR[Ds + i];
                    const float r4 = c.CR[Dd + i + s].y;
                    float4 v = c.VC[cc];
                    uint8_t *bp = c.bp + cc * 6;
                    int a0 = bp[2], a1 = bp[3], a2 = bp[4], a3 = bp[5];
                    amax1(v.x, a0, __fadd_rn(l3, i3.x), w - s);
                    amax1(v.y, a1, __fadd_rn(l3, i3.y), w - s);
                    amax1(v.z, a2, __fadd_rn(i4.x, r4), s - 1);
                    amax1(v.w, a3, __fadd_rn(i4.y, r4), s - 1);
                    bp[2] = (uint8_t)a0; bp[3] = (uint8_t)a1; bp[4] = (uint8_t)a2; bp[5] = (uint8_t)a3;
                    if (w == s) {
                        if (i == 0 && w != len) { v.z = p.mask_zero; v.w = p.mask_zero; }
                        c.CL[cc] = make_float2(v.x, v.y);
                        c.CR[cc] = make_float2(v.z, v.w);
                    } else {
                        c.VC[cc] = v;
                    }
                }
                blk_sync<NT>();
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
                        const int De = dbase(w - 1 - s, Nb);
                        const float2 la = c.CR[Ds + i], ra = c.CL[De + i + s + 1];
                        float2 v = c.VX[cc];
                        uint8_t *bp = c.bp + cc * 6;
                        int a0 = bp[0], a1 = bp[1];
                        amax1(v.x, a0, __fadd_rn(la.y, ra.x), s);  // split r - i = s
                        amax1(v.y, a1, __fadd_rn(la.x, ra.y), s);
                        if (w - 1 - s != s) {
                            const float2 lb = c.CR[De + i], rb = c.CL[Ds + j - s];
                            amax1(v.x, a0, __fadd_rn(lb.y, rb.x), w - 1 - s);
                            amax1(v.y, a1, __fadd_rn(lb.x, rb.y), w - 1 - s);
                        }
                        bp[0] = (uint8_t)a0; bp[1] = (uint8_t)a1;
                        if (w == s + 1) {
                            const float2 arcl = c.IL[cc], arcr = c.IR[cc];
                            c.IL[cc] = make_float2(__fadd_rn(v.x, arcl.x), __fadd_rn(v.x, arcl.y));
                            c.IR[cc] = make_float2(__fadd_rn(v.y, arcr.x), __fadd_rn(v.y, arcr.y));
                        } else {
                            c.VX[cc] = v;
                        }
                    } else {
                        const int cc = i0 + (t - nI);
                        const int w = cw[cc] >> 8, i = cw[cc] & 255, j = i + w;
                        const int Dd = dbase(w - s, Nb);
                        const float l3 = c.CL[Ds + i].y;
                        const float2 i3 = c.IL[Dd + i + s];
                        const float2 i4 = c.IR[Dd + i];
                        const float r4 = c.CR[Ds + j - s].y;
                        float4 v = c.VC[cc];
                        uint8_t *bp = c.bp + cc * 6;
                        int a0 = bp[2], a1 = bp[3], a2 = bp[4], a3 = bp[5];
                        amax1(v.x, a0, __fadd_rn(l3, i3.x), s);          // CL split r - i = s
                        amax1(v.y, a1, __fadd_rn(l3, i3.y), s);
                        amax1(v.z, a2, __fadd_rn(i4.x, r4), w - s - 1);  // CR split r - i - 1, r = j - s
                        amax1(v.w, a3, __fadd_rn(i4.y, r4), w - s - 1);
                        bp[2] = (uint8_t)a0; bp[3] = (uint8_t)a1; bp[4] = (uint8_t)a2; bp[5] = (uint8_t)a3;
                        c.VC[cc] = v;
                    }
                }
                blk_sync<NT>();
            }
        }
    }
    if (prof) p.prof[5] = clock64() - t0c;
    if (tid == 0) p.best[b] = c.CR[cidx(0, len, Nb)].y;

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
    blk_sync<NT>();
    if (prof) p.prof[6] = clock64() - t0c;
}

// ---------------------------------------------------------------------------------------------
// kernel: persistent CTAs stride over (sentence, semiring) work items 
// ---------------------------------------------------------------------------------------------
template <int NT, int CPT, bool GC = false>
__global__ void __launch_bounds__(NT, NT == 512 ? (CPT >= 4 ? 1 : 2) : (NT == 256 ? 3 : (NT == 128 ? 8 : (NT == 64 ? 16 : 1)))) dmv_frontier_kernel(DmvArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int total = p.B * p.npass;
    const int nsm = p.nsm;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int item = t;
        // (not with the host-memory hand-off: there a max CTA spins on its sentence's log CTA, which must have the
        // lower block index so that it is scheduled first whatever else shares the GPU)
        if (total <= (int)gridDim.x && t >= nsm && !p.share) item = total - 1 - (t - nsm);
        int b, which;
        if (p.npass == 2) { which = item >= p.B; b = which ? item - p.B : item; }
        else { which = p.first_pass; b = item; }
        if (p.only && p.only[b] == 0) continue;  // follow-up launch of the gather schedule: flagged sentences only
        const int len = clamp_len(p, b);  // read once: the lengths may live in host memory (one PCIe round trip)
        if (len + 1 < p.nb_lo || len + 1 > p.nb_hi) continue;
        unsigned char *chart = GC ? reinterpret_cast<unsigned char *>(p.workspace) + (size_t)blockIdx.x * p.ws_stride : nullptr;
#ifdef VLGAE_TIMELINE  // debug build (nvcc -DVLGAE_TIMELINE): the pointer kept live across the sweeps costs the 64-register
                       // variants 2 us on the cfg2 batch, so the default build does not carry it
        long long *tl = (p.prof && p.prof_all && threadIdx.x == 0) ? p.prof + 8 + ((size_t)which * p.B + b) * 2 : nullptr;
        if (tl) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl[0]));
#endif
        if (which == 0) log_pass<NT, CPT, GC>(p, b, len, smem_raw, chart);
        else max_pass<NT, CPT, GC>(p, b, len, smem_raw, chart);
#ifdef VLGAE_TIMELINE
        if (tl) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl[1]));
#endif
    }
}

// Warp-per-sentence variant for short sentences in the throughput regime: WPC sentences per CTA, each warp runs the
// same sweeps with __syncwarp as its barrier (a phase of a 12-word chart has < 32 active cells; a block barrier and 3 idle
// warps per phase cost more than the work), state in registers (<= CPT cells per lane).
template <int WPC, int CPT>
__global__ void __launch_bounds__(32 * WPC, CPT <= 3 ? 6 : 4) dmv_frontier_warp_kernel(DmvArgs p, int slice_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    unsigned char *slice = smem_raw + (size_t)warp * slice_bytes;
    const int total = p.B * p.npass;
    for (int item = blockIdx.x * WPC + warp; item < total; item += gridDim.x * WPC) {
        int b, which;
        if (p.npass == 2) { which = item >= p.B; b = which ? item - p.B : item; }
        else { which = p.first_pass; b = item; }
        if (p.only && p.only[b] == 0) continue;
        const int len = clamp_len(p, b);
        if (len + 1 < p.nb_lo || len + 1 > p.nb_hi) continue;
        if (which == 0) log_pass<32, CPT, false>(p, b, len, slice, nullptr);
        else max_pass<32, CPT, false>(p, b, len, slice, nullptr);
    }
}

size_t frontier_small_bytes(int cap) { return small_bytes(cap); }  // dec + cell table + offsets: always shared memory
size_t frontier_chart_only_bytes(int cap, int passes, bool reg_state) {
    const size_t nc = ncells(cap);
    size_t s = 0;
    if (passes & 1) s = nc * (reg_state ? 72 : 88) + 16;
    if (passes & 2) {
        const size_t m = nc * 56 + (size_t)(2 * (2 * cap + 2)) * 4 + nc * 6 + 16;
        s = m > s ? m : s;
    }
    return s;
}
size_t frontier_bytes(int cap, int passes, bool reg_state) {
    return (frontier_small_bytes(cap) + frontier_chart_only_bytes(cap, passes, reg_state) + 255) & ~(size_t)255;
}

}  // namespace

bool dmv_frontier_fits(int cap, int passes, int smem_optin) { return cap <= 256 && frontier_bytes(cap, passes, false) <= (size_t)smem_optin; }
size_t dmv_frontier_chart_bytes(int N, int passes) { return (frontier_chart_only_bytes(N, passes, false) + 255) & ~(size_t)255; }

cudaError_t launch_dmv_frontier(DmvArgs a, int passes, int cap, int threads, bool reg_state, int sm_count, int max_grid,
                                cudaStream_t st) {
    // a.workspace != null: the chart lives in the CTA's slice of the workspace (stride a.ws_stride, max_grid slices)
    const int total = a.B * a.npass;
    const bool global_chart = a.workspace != nullptr;
    if (global_chart) reg_state = false;
    a.smem_n = cap;
    auto go = [&](auto kern, int nt, bool regs) -> cudaError_t {
        const size_t smem = global_chart ? ((frontier_small_bytes(cap) + 255) & ~(size_t)255) : frontier_bytes(cap, passes, regs);
        // the attribute / occupancy queries cost ~10 us of host time per call: remember them per (variant, shared memory)
        struct Cached { const void *fn; size_t smem; int occ, dev; };
        static thread_local Cached cache[32];
        static thread_local int ncache = 0;
        int occ = 0, dev = 0;
        cudaGetDevice(&dev);
        for (int k = 0; k < ncache; ++k)
            if (cache[k].fn == (const void *)kern && cache[k].smem == smem && cache[k].dev == dev) occ = cache[k].occ;
        if (occ == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem);
            if (e != cudaSuccess) return e;
            if (occ < 1) occ = 1;
            // the attribute is sticky and only ever has to grow: keep one entry per kernel at its largest size
            int slot = -1;
            for (int k = 0; k < ncache; ++k) if (cache[k].fn == (const void *)kern && cache[k].dev == dev) slot = k;
            if (slot < 0 && ncache < 32) slot = ncache++;
            if (slot >= 0) cache[slot] = Cached{(const void *)kern, smem, occ, dev};
        }
        int grid = sm_count * occ;
        if (grid > total) grid = total;
        if (global_chart && grid > max_grid) grid = max_grid;
        kern<<<grid, nt, smem, st>>>(a);
        return cudaGetLastError();
    };
    // running state in registers (thread owns <= 4 target cells for the whole sweep) or in shared memory (any size)
    static const int env_smem_acc = [] { const char *v = getenv("VLGAE_FRONTIER_SMEM_ACC"); return v && *v ? atoi(v) : -1; }();
    if (env_smem_acc >= 0) reg_state = env_smem_acc == 0;
    const int cells = reg_state ? ncells(cap) - cap : (1 << 30);
    if (threads == 32 && !global_chart) {  // warp per sentence, 4 sentences per CTA
        constexpr int WPC = 4;
        const int cells = ncells(cap) - cap;
        if (cells <= 5 * 32) {
            const size_t slice = (frontier_bytes(cap, passes, true) + 15) & ~(size_t)15;
            auto gow = [&](auto kern) -> cudaError_t {
                const size_t smem = slice * WPC;
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                int occ = 0;
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * WPC, smem);
                if (e != cudaSuccess) return e;
                if (occ < 1) occ = 1;
                int grid = sm_count * occ, need = (total + WPC - 1) / WPC;
                if (grid > need) grid = need;
                kern<<<grid, 32 * WPC, smem, st>>>(a, (int)slice);
                return cudaGetLastError();
            };
            return cells <= 3 * 32 ? gow(dmv_frontier_warp_kernel<WPC, 3>) : gow(dmv_frontier_warp_kernel<WPC, 5>);
        }
        threads = 64;
    }
    if (global_chart) return threads <= 512 ? go(dmv_frontier_kernel<512, 0, true>, 512, false) : go(dmv_frontier_kernel<1024, 0, true>, 1024, false);
    if (threads <= 64) return cells <= 128 ? go(dmv_frontier_kernel<64, 2>, 64, true) : go(dmv_frontier_kernel<64, 0>, 64, false);
    if (threads <= 128) return cells <= 256 ? go(dmv_frontier_kernel<128, 2>, 128, true) : go(dmv_frontier_kernel<128, 0>, 128, false);
    if (threads <= 256) {
        if (cells <= 512) return go(dmv_frontier_kernel<256, 2>, 256, true);
        return cells <= 1024 ? go(dmv_frontier_kernel<256, 4>, 256, true) : go(dmv_frontier_kernel<256, 0>, 256, false);
    }
    if (threads <= 512) {
        if (cells <= 512) return go(dmv_frontier_kernel<512, 1>, 512, true);
        if (cells <= 1024) return go(dmv_frontier_kernel<512, 2>, 512, true);
        // charts of 46 .. 72 positions: five cells per thread in registers (one CTA per SM either way)
        return cells <= 2560 ? go(dmv_frontier_kernel<512, 5>, 512, true) : go(dmv_frontier_kernel<512, 0>, 512, false);
    }
    return cells <= 1024 ? go(dmv_frontier_kernel<1024, 1>, 1024, true) : go(dmv_frontier_kernel<1024, 0>, 1024, false);
}

}  // namespace vlgae
