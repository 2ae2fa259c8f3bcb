// dmv_kernels.cu -- DMV chart DP for sm_100a: one CTA owns one sentence at a time.
//
// Replaces the reference's O(N)-launch, autograd-through-the-chart path
//   /root/reference/src/model/torch_struct/dmv.py:19-66      (DMV1oStruct._dp)
//   /root/reference/src/model/torch_struct/helpers.py:118-154 (marginals / argmax by autograd)
// with three sweeps that never leave the SM:
//   inside  (log semiring)  -> Z
//   outside (explicit reverse sweep, no autograd) -> d Z / d attach (arc marginals), d Z / d dec
//   Viterbi (max semiring, first-max back-pointers) -> best score, heads, arc indicator, decision counts
//
// Chart storage ("diagonal-major"): the item spanning positions lo..hi (d = hi - lo) lives at
//   cidx(lo, d) = d * Nb - d (d - 1) / 2 + lo,        Nb = len + 1 positions incl. ROOT,
// so that for a fixed width the cells of neighbouring spans are contiguous (bank-conflict-free
// when consecutive lanes own consecutive spans).  Both valences of an item are one float2
// (.x = HASCHILD, .y = NOCHILD).  Per cell the log pass keeps 10 float2 (80 B), the max pass
// 4 float2 + 6 back-pointer bytes.  For Nb <= ~80 everything is in shared memory; longer
// sentences use the same code with the chart in a per-CTA global workspace (L2-resident).
//
// Work decomposition for width w: cell (i, j = i + w) is owned by a group of g lanes (g a power
// of two <= 32 chosen per width so that (Nb - w) * g fills the CTA); lane `sub` handles the split
// points r' = sub, sub + g, ...  A group computes the two incomplete items of its span and then,
// without a block barrier, its two complete items (the only width-w operands those need are the
// group's own), so there is ONE __syncthreads per width.  The reverse sweep keeps the
// contributions of complete-item parents and incomplete-item parents in separate accumulators,
// which makes every read-modify-write target unique within a width: no atomics, one barrier.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmv_kernels.cuh"

namespace vlgae {

namespace {

constexpr float NEG_BIG = -3.0e38f;  // finite stand-in for -inf (no NaN from (-inf) - (-inf))
constexpr int KCH = 4;               // split points per lane per chunk of the streaming logsumexp

__device__ __forceinline__ int cidx(int lo, int d, int Nb) { return d * Nb - ((d * (d - 1)) >> 1) + lo; }

__device__ __forceinline__ unsigned group_mask(int g) {
    const int lane = threadIdx.x & 31;
    return g >= 32 ? 0xffffffffu : (((1u << g) - 1u) << (lane & ~(g - 1)));
}

// lanes per cell for width w with ncell cells: fill the CTA, never more lanes than split points (rounded up)
__device__ __forceinline__ int lanes_per_cell(int ncell, int w, int nthreads) {
    int g = 1;
    while (g < 32 && g < w && ncell * (g << 1) <= nthreads) g <<= 1;
    return g;
}

// streaming logsumexp state: value = m + log(s)
struct Lse {
    float m, s;
    __device__ __forceinline__ void init() { m = NEG_BIG; s = 0.f; }
    __device__ __forceinline__ void add_chunk(const float (&t)[KCH]) {
        float cm = t[0];
#pragma unroll
        for (int k = 1; k < KCH; ++k) cm = fmaxf(cm, t[k]);
        const float nm = fmaxf(m, cm);
        s *= __expf(m - nm);
        m = nm;
#pragma unroll
        for (int k = 0; k < KCH; ++k) s += __expf(t[k] - nm);
    }
    // combine across the g lanes of a group (all lanes end with the same value)
    __device__ __forceinline__ float finish(int g, unsigned mask) {
        float gm = m;
        for (int o = g >> 1; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(mask, gm, o));
        float gs = s * __expf(m - gm);
        for (int o = g >> 1; o > 0; o >>= 1) gs += __shfl_xor_sync(mask, gs, o);
        return gm + __logf(gs);
    }
};

// first-max state: value and smallest split index attaining it (torch.max tie rule)
struct ArgMax {
    float v;
    int a;
    __device__ __forceinline__ void init() { v = NEG_BIG; a = 0x7fffffff; }
    __device__ __forceinline__ void add(float t, int r) {
        if (t > v) { v = t; a = r; }  // r increases within a lane: strict > keeps the first
    }
    __device__ __forceinline__ void finish(int g, unsigned mask) {
        for (int o = g >> 1; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(mask, v, o);
            const int oa = __shfl_xor_sync(mask, a, o);
            if (ov > v || (ov == v && oa < a)) { v = ov; a = oa; }
        }
    }
};

struct LogChart {
    float2 *CL, *CR, *IL, *IR, *X;     // inside values; X = (XL, XR) pre-arc reductions
    float2 *gCL, *gCR, *gCa, *gIL, *gIR;  // reverse sweep; gCa = (A-part of gCL[NO], A-part of gCR[NO])
};

__host__ __device__ inline int ncells(int Nb) { return Nb * (Nb + 1) / 2; }

// ---------------------------------------------------------------------------------------------
// log semiring: inside + outside for one sentence
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ void log_pass(const DmvArgs &p, int b, float2 *mem, float *sdec) {
    const int tid = threadIdx.x;
    const int N = p.N;
    int len = (int)p.lengths[b];
    len = len < 0 ? 0 : (len > N - 1 ? N - 1 : len);
    const int Nb = len + 1;
    const int nc = ncells(Nb);
    LogChart c;
    c.CL = mem; c.CR = mem + nc; c.IL = mem + 2 * nc; c.IR = mem + 3 * nc; c.X = mem + 4 * nc;
    c.gCL = mem + 5 * nc; c.gCR = mem + 6 * nc; c.gCa = mem + 7 * nc; c.gIL = mem + 8 * nc; c.gIR = mem + 9 * nc;
    const bool want_grad = (p.gdec != nullptr) || (p.gattach != nullptr);

    const float *dec = p.dec + (size_t)b * N * 8;
    const float *attach = p.attach + (size_t)b * N * N * 2;
    for (int t = tid; t < Nb * 8; t += NT) sdec[t] = dec[t];
    if (want_grad)
        for (int t = tid; t < 5 * nc; t += NT) c.gCL[t] = make_float2(0.f, 0.f);
    __syncthreads();
    // width-0 complete items = STOP decisions (dmv.py:39-40); arc scores pre-added into the I cells
    // (attach + dec[GO] is formed first, exactly as dmv.py:36-37 does)
    for (int i = tid; i < Nb; i += NT) {
        c.CL[i] = make_float2(sdec[i * 8 + 0 * 4 + 0 * 2 + 1], sdec[i * 8 + 0 * 4 + 1 * 2 + 1]);
        c.CR[i] = make_float2(sdec[i * 8 + 1 * 4 + 0 * 2 + 1], sdec[i * 8 + 1 * 4 + 1 * 2 + 1]);
    }
    for (int t = tid; t < Nb * Nb; t += NT) {
        const int h = t / Nb, ch = t - h * Nb;
        if (h == ch) continue;
        const float2 a = *reinterpret_cast<const float2 *>(attach + ((size_t)h * N + ch) * 2);
        if (ch < h)
            c.IL[cidx(ch, h - ch, Nb)] = make_float2(a.x + sdec[h * 8 + 0 * 4 + 0 * 2 + 0], a.y + sdec[h * 8 + 0 * 4 + 1 * 2 + 0]);
        else
            c.IR[cidx(h, ch - h, Nb)] = make_float2(a.x + sdec[h * 8 + 1 * 4 + 0 * 2 + 0], a.y + sdec[h * 8 + 1 * 4 + 1 * 2 + 0]);
    }
    __syncthreads();

    // ---------------- inside ----------------
    for (int w = 1; w < Nb; ++w) {
        const int ncell = Nb - w;
        const int g = lanes_per_cell(ncell, w, NT);
        const unsigned mask = group_mask(g);
        const int sub = tid & (g - 1);
        for (int i = tid / g; i < ncell; i += NT / g) {
            const int own = cidx(i, w, Nb);
            Lse xl, xr;
            xl.init(); xr.init();
            for (int r0 = sub; r0 < w; r0 += g * KCH) {
                float tl[KCH], tr[KCH];
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const int rp = r0 + k * g;
                    if (rp < w) {
                        const float2 a = c.CR[cidx(i, rp, Nb)];                   // CR[i][i+rp]
                        const float2 bb = c.CL[cidx(i + rp + 1, w - 1 - rp, Nb)];  // CL[j][i+rp+1]
                        tl[k] = a.y + bb.x;                                        // step 1 (dmv.py:50)
                        tr[k] = a.x + bb.y;                                        // step 2 (dmv.py:54)
                    } else {
                        tl[k] = NEG_BIG; tr[k] = NEG_BIG;
                    }
                }
                xl.add_chunk(tl); xr.add_chunk(tr);
            }
            const float XL = xl.finish(g, mask), XR = xr.finish(g, mask);
            const float2 arcL = c.IL[own], arcR = c.IR[own];
            const float2 il = make_float2(XL + arcL.x, XL + arcL.y);  // dmv.py:51-52
            const float2 ir = make_float2(XR + arcR.x, XR + arcR.y);  // dmv.py:55-56
            Lse l0, l1, q0, q1;
            l0.init(); l1.init(); q0.init(); q1.init();
            for (int r0 = sub; r0 < w; r0 += g * KCH) {
                float t0[KCH], t1[KCH], u0[KCH], u1[KCH];
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const int rp = r0 + k * g;
                    if (rp < w) {
                        const float cl = c.CL[cidx(i, rp, Nb)].y;                                // CL[i+rp][i][NO]
                        const float2 e = rp == 0 ? il : c.IL[cidx(i + rp, w - rp, Nb)];           // IL[j][i+rp]
                        t0[k] = cl + e.x; t1[k] = cl + e.y;                                       // step 3 (dmv.py:58)
                        const float2 f = rp == w - 1 ? ir : c.IR[cidx(i, rp + 1, Nb)];            // IR[i][i+1+rp]
                        const float cr = c.CR[cidx(i + 1 + rp, w - 1 - rp, Nb)].y;                // CR[i+1+rp][j][NO]
                        u0[k] = f.x + cr; u1[k] = f.y + cr;                                       // step 4 (dmv.py:61)
                    } else {
                        t0[k] = NEG_BIG; t1[k] = NEG_BIG; u0[k] = NEG_BIG; u1[k] = NEG_BIG;
                    }
                }
                l0.add_chunk(t0); l1.add_chunk(t1); q0.add_chunk(u0); q1.add_chunk(u1);
            }
            float2 cl2 = make_float2(l0.finish(g, mask), l1.finish(g, mask));
            float2 cr2 = make_float2(q0.finish(g, mask), q1.finish(g, mask));
            if (i == 0 && w != len) cr2 = make_float2(p.mask_zero, p.mask_zero);  // single root (dmv.py:63)
            __syncwarp(mask);  // every lane has read the pre-added arc scores of `own`
            if (sub == 0) {
                c.IL[own] = il; c.IR[own] = ir; c.X[own] = make_float2(XL, XR);
                c.CL[own] = cl2; c.CR[own] = cr2;
            }
        }
        __syncthreads();
    }
    if (tid == 0) p.Z[b] = c.CR[cidx(0, len, Nb)].y;  // dmv.py:65
    if (!want_grad) { __syncthreads(); return; }

    // ---------------- outside (reverse sweep; replaces helpers.py:150-154) ----------------
    if (tid == 0) c.gCa[cidx(0, len, Nb)].y = p.gZ ? p.gZ[b] : 1.f;
    __syncthreads();
    for (int w = Nb - 1; w >= 1; --w) {
        const int ncell = Nb - w;
        const int g = lanes_per_cell(ncell, w, NT);
        const unsigned mask = group_mask(g);
        const int sub = tid & (g - 1);
        for (int i = tid / g; i < ncell; i += NT / g) {
            const int own = cidx(i, w, Nb);
            // A: complete-item parents (steps 4 and 3 transposed)
            float2 gcr = c.gCR[own], gcl = c.gCL[own];
            const float2 ga = c.gCa[own];
            gcr.y += ga.y; gcl.y += ga.x;
            if (i == 0 && w != len) gcr = make_float2(0.f, 0.f);  // masked cell passes nothing back
            const float2 outR = c.CR[own], outL = c.CL[own];
            for (int rp = sub; rp < w; rp += g) {
                {
                    const int ci = cidx(i, rp + 1, Nb), cc = cidx(i + 1 + rp, w - 1 - rp, Nb);
                    const float2 f = c.IR[ci];
                    const float h = c.CR[cc].y;
                    const float p0 = gcr.x * __expf(f.x + h - outR.x);
                    const float p1 = gcr.y * __expf(f.y + h - outR.y);
                    float2 t = c.gIR[ci]; t.x += p0; t.y += p1; c.gIR[ci] = t;
                    c.gCa[cc].y += p0 + p1;
                }
                {
                    const int cc = cidx(i, rp, Nb), ci = cidx(i + rp, w - rp, Nb);
                    const float cl = c.CL[cc].y;
                    const float2 e = c.IL[ci];
                    const float p0 = gcl.x * __expf(cl + e.x - outL.x);
                    const float p1 = gcl.y * __expf(cl + e.y - outL.y);
                    c.gCa[cc].x += p0 + p1;
                    float2 t = c.gIL[ci]; t.x += p0; t.y += p1; c.gIL[ci] = t;
                }
            }
            __syncwarp(mask);
            // B: incomplete-item parents (steps 2 and 1 transposed)
            const float2 giR = c.gIR[own], giL = c.gIL[own];
            const float gxR = giR.x + giR.y, gxL = giL.x + giL.y;
            const float2 X = c.X[own];
            for (int rp = sub; rp < w; rp += g) {
                const int ca = cidx(i, rp, Nb), cb = cidx(i + rp + 1, w - 1 - rp, Nb);
                const float2 a = c.CR[ca];
                const float2 bb = c.CL[cb];
                const float pL = gxL * __expf(a.y + bb.x - X.x);
                const float pR = gxR * __expf(a.x + bb.y - X.y);
                float2 t = c.gCR[ca]; t.x += pR; t.y += pL; c.gCR[ca] = t;
                float2 u = c.gCL[cb]; u.x += pL; u.y += pR; c.gCL[cb] = u;
            }
        }
        __syncthreads();
    }
    // ---------------- outputs ----------------
    if (p.gattach) {
        float2 *ga = reinterpret_cast<float2 *>(p.gattach + (size_t)b * N * N * 2);
        for (int t = tid; t < N * N; t += NT) {
            const int h = t / N, ch = t - h * N;
            float2 v = make_float2(0.f, 0.f);
            if (h < Nb && ch < Nb && h != ch) v = ch < h ? c.gIL[cidx(ch, h - ch, Nb)] : c.gIR[cidx(h, ch - h, Nb)];
            ga[t] = v;
        }
    }
    if (p.gdec) {
        float *gd = p.gdec + (size_t)b * N * 8;
        for (int t = tid; t < N * 2; t += NT) {
            const int i = t >> 1, dir = t & 1;
            float2 go = make_float2(0.f, 0.f), stop = make_float2(0.f, 0.f);
            if (i < Nb) {
                if (dir == 0) {
                    for (int ch = 0; ch < i; ++ch) { const float2 v = c.gIL[cidx(ch, i - ch, Nb)]; go.x += v.x; go.y += v.y; }
                    stop = c.gCL[i]; stop.y += c.gCa[i].x;
                } else {
                    for (int d = 1; d < Nb - i; ++d) { const float2 v = c.gIR[cidx(i, d, Nb)]; go.x += v.x; go.y += v.y; }
                    stop = c.gCR[i]; stop.y += c.gCa[i].y;
                }
            }
            // [dir][val][decision]
            *reinterpret_cast<float4 *>(gd + i * 8 + dir * 4) = make_float4(go.x, stop.x, go.y, stop.y);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// max semiring: Viterbi chart with first-max back-pointers + parallel back-trace
// ---------------------------------------------------------------------------------------------
struct MaxChart {
    float2 *CL, *CR, *IL, *IR;
    uint32_t *bpA;  // XL | XR << 8 | CL[HAS] << 16 | CL[NO] << 24
    uint16_t *bpB;  // CR[HAS] | CR[NO] << 8
};

// items of the back-trace: kind (0 CR, 1 CL, 2 IR, 3 IL) | v << 2 | lo << 3 | hi << 12
__device__ __forceinline__ int mk_item(int kind, int v, int lo, int hi) { return kind | (v << 2) | (lo << 3) | (hi << 12); }

template <int NT>
__device__ void max_pass(const DmvArgs &p, int b, float2 *mem, float *sdec) {
    const int tid = threadIdx.x;
    const int N = p.N;
    int len = (int)p.lengths[b];
    len = len < 0 ? 0 : (len > N - 1 ? N - 1 : len);
    const int Nb = len + 1;
    const int nc = ncells(Nb);
    MaxChart c;
    c.CL = mem; c.CR = mem + nc; c.IL = mem + 2 * nc; c.IR = mem + 3 * nc;
    c.bpA = reinterpret_cast<uint32_t *>(mem + 4 * nc);
    c.bpB = reinterpret_cast<uint16_t *>(c.bpA + nc);
    int *queue = reinterpret_cast<int *>(c.bpB + ((nc + 1) & ~1));  // 2 x (2 Nb + 2) ints

    const float *dec = p.dec + (size_t)b * N * 8;
    const float *attach = p.attach + (size_t)b * N * N * 2;
    for (int t = tid; t < Nb * 8; t += NT) sdec[t] = dec[t];
    __syncthreads();
    for (int i = tid; i < Nb; i += NT) {
        c.CL[i] = make_float2(sdec[i * 8 + 0 * 4 + 0 * 2 + 1], sdec[i * 8 + 0 * 4 + 1 * 2 + 1]);
        c.CR[i] = make_float2(sdec[i * 8 + 1 * 4 + 0 * 2 + 1], sdec[i * 8 + 1 * 4 + 1 * 2 + 1]);
    }
    for (int t = tid; t < Nb * Nb; t += NT) {
        const int h = t / Nb, ch = t - h * Nb;
        if (h == ch) continue;
        const float2 a = *reinterpret_cast<const float2 *>(attach + ((size_t)h * N + ch) * 2);
        if (ch < h)
            c.IL[cidx(ch, h - ch, Nb)] = make_float2(__fadd_rn(a.x, sdec[h * 8 + 0]), __fadd_rn(a.y, sdec[h * 8 + 2]));
        else
            c.IR[cidx(h, ch - h, Nb)] = make_float2(__fadd_rn(a.x, sdec[h * 8 + 4]), __fadd_rn(a.y, sdec[h * 8 + 6]));
    }
    // outputs that the back-trace only dots with ones are zero-filled up front
    if (p.arcs) {
        float2 *z = reinterpret_cast<float2 *>(p.arcs + (size_t)b * N * N * 2);
        for (int t = tid; t < N * N; t += NT) z[t] = make_float2(0.f, 0.f);
    }
    if (p.vgdec) for (int t = tid; t < N * 8; t += NT) p.vgdec[(size_t)b * N * 8 + t] = 0.f;
    if (p.heads) for (int t = tid; t < N; t += NT) p.heads[(size_t)b * N + t] = 0;
    __syncthreads();

    for (int w = 1; w < Nb; ++w) {
        const int ncell = Nb - w;
        const int g = lanes_per_cell(ncell, w, NT);
        const unsigned mask = group_mask(g);
        const int sub = tid & (g - 1);
        for (int i = tid / g; i < ncell; i += NT / g) {
            const int own = cidx(i, w, Nb);
            ArgMax xl, xr;
            xl.init(); xr.init();
            for (int rp = sub; rp < w; rp += g) {
                const float2 a = c.CR[cidx(i, rp, Nb)];
                const float2 bb = c.CL[cidx(i + rp + 1, w - 1 - rp, Nb)];
                xl.add(__fadd_rn(a.y, bb.x), rp);
                xr.add(__fadd_rn(a.x, bb.y), rp);
            }
            xl.finish(g, mask); xr.finish(g, mask);
            const float2 arcL = c.IL[own], arcR = c.IR[own];
            const float2 il = make_float2(__fadd_rn(xl.v, arcL.x), __fadd_rn(xl.v, arcL.y));
            const float2 ir = make_float2(__fadd_rn(xr.v, arcR.x), __fadd_rn(xr.v, arcR.y));
            ArgMax l0, l1, q0, q1;
            l0.init(); l1.init(); q0.init(); q1.init();
            for (int rp = sub; rp < w; rp += g) {
                const float cl = c.CL[cidx(i, rp, Nb)].y;
                const float2 e = rp == 0 ? il : c.IL[cidx(i + rp, w - rp, Nb)];
                l0.add(__fadd_rn(cl, e.x), rp); l1.add(__fadd_rn(cl, e.y), rp);
                const float2 f = rp == w - 1 ? ir : c.IR[cidx(i, rp + 1, Nb)];
                const float cr = c.CR[cidx(i + 1 + rp, w - 1 - rp, Nb)].y;
                q0.add(__fadd_rn(f.x, cr), rp); q1.add(__fadd_rn(f.y, cr), rp);
            }
            l0.finish(g, mask); l1.finish(g, mask); q0.finish(g, mask); q1.finish(g, mask);
            float2 cr2 = make_float2(q0.v, q1.v);
            if (i == 0 && w != len) cr2 = make_float2(p.mask_zero, p.mask_zero);
            __syncwarp(mask);
            if (sub == 0) {
                c.IL[own] = il; c.IR[own] = ir;
                c.CL[own] = make_float2(l0.v, l1.v); c.CR[own] = cr2;
                c.bpA[own] = (uint32_t)xl.a | ((uint32_t)xr.a << 8) | ((uint32_t)l0.a << 16) | ((uint32_t)l1.a << 24);
                c.bpB[own] = (uint16_t)((uint32_t)q0.a | ((uint32_t)q1.a << 8));
            }
        }
        __syncthreads();
    }
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
                        if (p.vgdec) p.vgdec[(size_t)b * N * 8 + lo * 8 + (kind == 0 ? 4 : 0) + v * 2 + 1] = 1.f;
                    } else {
                        const int own = cidx(lo, d, Nb);
                        const uint32_t ba = c.bpA[own];
                        const uint32_t bb = c.bpB[own];
                        if (kind == 0) {  // CR(lo,hi,v) -> IR(lo,r,v) + CR(r,hi,NO), r = lo+1+bp
                            const int r = lo + 1 + (int)((bb >> (8 * v)) & 255);
                            c1 = mk_item(2, v, lo, r); c2 = mk_item(0, 1, r, hi);
                        } else if (kind == 1) {  // CL(hi,lo,v) -> CL(r,lo,NO) + IL(hi,r,v), r = lo+bp
                            const int r = lo + (int)((ba >> (16 + 8 * v)) & 255);
                            c1 = mk_item(1, 1, lo, r); c2 = mk_item(3, v, r, hi);
                        } else if (kind == 2) {  // IR: arc lo -> hi; XR -> CR(lo,r,HAS) + CL(hi,r+1,NO)
                            const int r = lo + (int)((ba >> 8) & 255);
                            c1 = mk_item(0, 0, lo, r); c2 = mk_item(1, 1, r + 1, hi);
                            if (p.heads) p.heads[(size_t)b * N + hi] = lo;
                            if (p.arcs) p.arcs[(((size_t)b * N + lo) * N + hi) * 2 + v] = 1.f;
                            if (p.vgdec) p.vgdec[(size_t)b * N * 8 + lo * 8 + 4 + v * 2 + 0] = 1.f;
                        } else {  // IL: arc hi -> lo; XL -> CR(lo,r,NO) + CL(hi,r+1,HAS)
                            const int r = lo + (int)(ba & 255);
                            c1 = mk_item(0, 1, lo, r); c2 = mk_item(1, 0, r + 1, hi);
                            if (p.heads) p.heads[(size_t)b * N + lo] = hi;
                            if (p.arcs) p.arcs[(((size_t)b * N + hi) * N + lo) * 2 + v] = 1.f;
                            if (p.vgdec) p.vgdec[(size_t)b * N * 8 + hi * 8 + 0 + v * 2 + 0] = 1.f;
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
}

// ---------------------------------------------------------------------------------------------
// kernel: persistent CTAs stride over (sentence, semiring) work items
// ---------------------------------------------------------------------------------------------
template <int NT, bool SMEM>
__global__ void __launch_bounds__(NT) dmv_kernel(DmvArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sdec = reinterpret_cast<float *>(smem_raw);
    float2 *mem;
    if (SMEM)
        mem = reinterpret_cast<float2 *>(smem_raw + (((size_t)p.N * 8 * sizeof(float) + 15) & ~(size_t)15));
    else
        mem = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(p.workspace) + (size_t)blockIdx.x * p.ws_stride);
    const int total = p.B * p.npass;
    // static round-robin over (sentence, semiring) work items; batches arrive sorted by length
    // (reference sampler.py:135-136), so consecutive items cost about the same
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / p.npass;
        const int which = p.npass == 2 ? (t & 1) : p.first_pass;
        if (which == 0) log_pass<NT>(p, b, mem, sdec);
        else max_pass<NT>(p, b, mem, sdec);
    }
}

__global__ void merge_kernel(const float *dec, const float *attach, const float *root, int B, int n, float one,
                             float zero, float *dec_w, float *attach_w) {
    // distributions.py:253-265
    const int N = n + 1;
    const size_t na = (size_t)B * N * N * 2, nd = (size_t)B * N * 8;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < na + nd; t += (size_t)gridDim.x * blockDim.x) {
        if (t < na) {
            const int v = t & 1;
            size_t r = t >> 1;
            const int c = r % N; r /= N;
            const int h = r % N;
            const size_t b = r / N;
            float x = zero;
            if (h == 0) { if (c >= 1 && v == 1) x = root[b * n + (c - 1)]; }
            else if (c >= 1) x = attach[((b * n + (h - 1)) * n + (c - 1)) * 2 + v];
            attach_w[t] = x;
        } else {
            const size_t u = t - na;
            const int k = u & 7;  // dir*4 + val*2 + decision
            const size_t r = u >> 3;
            const int i = r % N;
            const size_t b = r / N;
            float x;
            if (i == 0) x = (k >> 2) == 1 ? one : zero;
            else x = dec[(b * n + (i - 1)) * 8 + k];
            dec_w[u] = x;
        }
    }
}

__global__ void scale_rows_kernel(const float *in, const float *g, int B, size_t inner, float *out) {
    const size_t total = (size_t)B * inner;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
        out[t] = in[t] * g[t / inner];
}

__global__ void mufu_bench_kernel(int iters, float *sink) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    float e = a + 0.4f, f = a + 0.5f, g = a + 0.6f, h = a + 0.7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(h));
        }
    }
    if (a + b + c + d + e + f + g + h == 123.456f) sink[0] = a;
}

__global__ void fp32_bench_kernel(int iters, float *sink) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    float e = a + 0.4f, f = a + 0.5f, g = a + 0.6f, h = a + 0.7f;
    const float k = 1.0000001f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(k)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b) : "f"(k));
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(c) : "f"(k)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(d) : "f"(k));
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(e) : "f"(k)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(k));
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g) : "f"(k)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(h) : "f"(k));
        }
    }
    if (a + b + c + d + e + f + g + h == 123.456f) sink[0] = a;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host-side launch logic
// ---------------------------------------------------------------------------------------------
size_t log_chart_bytes(int N) { return (size_t)ncells(N) * 10 * sizeof(float2); }
size_t max_chart_bytes(int N) {
    const size_t nc = ncells(N);
    return nc * 4 * sizeof(float2) + nc * 4 + ((nc + 1) & ~(size_t)1) * 2 + (size_t)(2 * (2 * N + 2)) * 4 + 16;
}
static size_t dec_bytes(int N) { return ((size_t)N * 8 * sizeof(float) + 15) & ~(size_t)15; }

static int g_sm_count = 0, g_smem_optin = 0;
static cudaError_t device_info() {
    if (g_sm_count) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    return cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}

size_t dmv_chart_bytes(int N, int passes /*1 log, 2 max, 3 both*/) {
    size_t s = 0;
    if (passes & 1) s = log_chart_bytes(N);
    if (passes & 2) { const size_t m = max_chart_bytes(N); s = m > s ? m : s; }
    return (s + 255) & ~(size_t)255;
}

bool dmv_fits_smem(int N, int passes) {
    if (device_info() != cudaSuccess) return false;
    return dec_bytes(N) + dmv_chart_bytes(N, passes) <= (size_t)g_smem_optin;
}

int dmv_grid_for_workspace(int B) {
    if (device_info() != cudaSuccess) return 0;
    const int cap = g_sm_count * 4;
    return B * 2 < cap ? B * 2 : cap;
}

template <int NT>
static cudaError_t launch_nt(DmvArgs a, int passes, cudaStream_t st) {
    cudaError_t e = device_info();
    if (e != cudaSuccess) return e;
    const size_t chart = dmv_chart_bytes(a.N, passes);
    const size_t smem_need = dec_bytes(a.N) + chart;
    const int total = a.B * a.npass;
    if (smem_need <= (size_t)g_smem_optin) {
        auto k = dmv_kernel<NT, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem_need);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        int grid = g_sm_count * occ;
        if (grid > total) grid = total;
        k<<<grid, NT, smem_need, st>>>(a);
    } else {
        auto k = dmv_kernel<NT, false>;
        int grid = dmv_grid_for_workspace(a.B);
        if (grid > total) grid = total;
        a.ws_stride = chart;
        k<<<grid, NT, dec_bytes(a.N), st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_dmv(const DmvArgs &a, int passes, cudaStream_t st) {
    // short charts: 128 threads (more CTAs per SM); long charts: 256
    if (a.N <= 48) return launch_nt<128>(a, passes, st);
    return launch_nt<256>(a, passes, st);
}

cudaError_t launch_merge(const float *dec, const float *attach, const float *root, int B, int n, float one, float zero,
                         float *dec_w, float *attach_w, cudaStream_t st) {
    const size_t total = (size_t)B * (n + 1) * ((size_t)(n + 1) * 2 + 8);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    merge_kernel<<<grid, 256, 0, st>>>(dec, attach, root, B, n, one, zero, dec_w, attach_w);
    return cudaGetLastError();
}

cudaError_t launch_scale_rows(const float *in, const float *g, int B, size_t inner, float *out, cudaStream_t st) {
    const size_t total = (size_t)B * inner;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    scale_rows_kernel<<<grid, 256, 0, st>>>(in, g, B, inner, out);
    return cudaGetLastError();
}

cudaError_t launch_microbench(int which, int iters, float *sink, int *grid_out, int *block_out, cudaStream_t st) {
    cudaError_t e = device_info();
    if (e != cudaSuccess) return e;
    const int grid = g_sm_count * 8, block = 256;
    if (which == 0) mufu_bench_kernel<<<grid, block, 0, st>>>(iters, sink);
    else fp32_bench_kernel<<<grid, block, 0, st>>>(iters, sink);
    *grid_out = grid; *block_out = block;
    return cudaGetLastError();
}

}  // namespace vlgae
