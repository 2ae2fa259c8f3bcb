/*
 * oracle/dmv_oracle.c -- CPU restatement of the reference's DMV chart DP.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vlgae_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.
 *
 * What it restates (all citations relative to /root/reference):
 *   - DMV1oStruct._dp                 src/model/torch_struct/dmv.py:19-66
 *   - LogSemiring / MaxSemiring       src/model/torch_struct/semirings/semirings.py:127-148,173-207
 *   - _Struct.marginals (autograd)    src/model/torch_struct/helpers.py:118-154
 *       restated as an explicit reverse sweep over the chart
 *   - DepTree._dp (+_check_potentials) src/model/torch_struct/deptree.py:25-76,146-162
 *
 * Parity pin: the reference ships no tests or golden vectors for this path
 * ("parity unpinned" by the reference's own suite).  This restatement is
 * pinned instead against the reference itself, imported live in the build
 * container by tests/golden/gen_golden.py, whose outputs are committed under
 * tests/golden/ and checked by tests/test_oracle.py.
 *
 * Storage convention here (not the reference's shifted (N+1)x(N+1) chart):
 *   CR[i][j][v]  complete,   head i, right end j   (ref: C[i, j+1, v])
 *   CL[j][i][v]  complete,   head j, left  end i   (ref: C[j, i,   v])
 *   IR[i][j][v]  incomplete, arc i->j              (ref: I[i, j+1, v])
 *   IL[j][i][v]  incomplete, arc j->i              (ref: I[j, i,   v])
 * each a dense [N][N][2] block.
 *
 * The file is compiled twice: -DREAL=float (mirrors the reference's fp32
 * arithmetic) and -DREAL=double (a high-precision truth used to size the
 * tolerances).  Inputs are always float32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* dmv.py:7-15 */
enum { HASCHILD = 0, NOCHILD = 1, LEFT = 0, RIGHT = 1, GO = 0, STOP = 1 };

#define DEC(p, d, v, k) dec[(((size_t)(p) * 2 + (d)) * 2 + (v)) * 2 + (k)]
#define AT(A, a, b, v) A[((size_t)(a) * CS + (b)) * 2 + (v)]
#define XI(a, b) ((size_t)(a) * CS + (b))

static REAL real_exp(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static REAL real_log(REAL x) { return sizeof(REAL) == 4 ? (REAL)logf((float)x) : (REAL)log((double)x); }

/* torch.logsumexp over k terms: max, sum(exp(x - max)), log, + max
 * (semirings.py:131-132).  All scores are finite sentinels, never -inf. */
static REAL lse(const REAL *x, int k) {
    REAL m = x[0];
    for (int t = 1; t < k; ++t)
        if (x[t] > m) m = x[t];
    REAL s = 0;
    for (int t = 0; t < k; ++t) s += real_exp(x[t] - m);
    return real_log(s) + m;
}

/* torch.max(dim): value and FIRST maximal index (semirings.py:200-202). */
static REAL first_max(const REAL *x, int k, int *arg) {
    REAL m = x[0];
    int a = 0;
    for (int t = 1; t < k; ++t)
        if (x[t] > m) { m = x[t]; a = t; }
    *arg = a;
    return m;
}

typedef struct {
    REAL *CL, *CR, *IL, *IR; /* values [N][N][2] */
    REAL *XL, *XR;           /* pre-arc reductions [N][N] (indexed like IL / IR) */
} chart_t;

static int chart_alloc(chart_t *c, int N) {
    size_t n2 = (size_t)N * N;
    c->CL = (REAL *)calloc(n2 * 2, sizeof(REAL));
    c->CR = (REAL *)calloc(n2 * 2, sizeof(REAL));
    c->IL = (REAL *)calloc(n2 * 2, sizeof(REAL));
    c->IR = (REAL *)calloc(n2 * 2, sizeof(REAL));
    c->XL = (REAL *)calloc(n2, sizeof(REAL));
    c->XR = (REAL *)calloc(n2, sizeof(REAL));
    return c->CL && c->CR && c->IL && c->IR && c->XL && c->XR;
}
static void chart_free(chart_t *c) {
    free(c->CL); free(c->CR); free(c->IL); free(c->IR); free(c->XL); free(c->XR);
}

/*
 * Forward chart for one sentence.  semiring: 0 = log, 1 = max.
 * N   : number of chart positions actually swept (Nfull, or len+1 when trimmed)
 * Ns  : row stride of the input tensors (Nfull)
 * bp_*: optional back-pointers (first maximal split) for the max semiring.
 */
static void forward_one(const float *dec, const float *attach_in, int Ns, int N, int len, REAL mask_zero,
                        int semiring, chart_t *c, int *bpXL, int *bpXR, int *bpCL, int *bpCR, REAL *scratch) {
#define ATT(h, cc, v) attach_in[((size_t)(h) * Ns + (cc)) * 2 + (v)]
    const int CS = N;
    REAL *CL = c->CL, *CR = c->CR, *IL = c->IL, *IR = c->IR;
    int arg = 0;
    /* dmv.py:39-40 -- width-0 complete items are the STOP decisions */
    for (int i = 0; i < N; ++i)
        for (int v = 0; v < 2; ++v) {
            AT(CL, i, i, v) = (REAL)DEC(i, LEFT, v, STOP);
            AT(CR, i, i, v) = (REAL)DEC(i, RIGHT, v, STOP);
        }
    for (int w = 1; w < N; ++w) {
        for (int i = 0; i + w < N; ++i) {
            int j = i + w;
            /* step 1, dmv.py:50-52 */
            for (int r = i; r < j; ++r) scratch[r - i] = AT(CR, i, r, NOCHILD) + AT(CL, j, r + 1, HASCHILD);
            REAL xl = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
            if (bpXL) bpXL[XI(j, i)] = arg;
            c->XL[XI(j, i)] = xl;
            for (int v = 0; v < 2; ++v) {
                /* dmv.py:36 -- attach_left is formed in fp32 before the add */
                float al = ATT(j, i, v) + DEC(j, LEFT, v, GO);
                AT(IL, j, i, v) = xl + (REAL)(sizeof(REAL) == 4 ? al : (double)ATT(j, i, v) + (double)DEC(j, LEFT, v, GO));
            }
            /* step 2, dmv.py:54-56 */
            for (int r = i; r < j; ++r) scratch[r - i] = AT(CR, i, r, HASCHILD) + AT(CL, j, r + 1, NOCHILD);
            REAL xr = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
            if (bpXR) bpXR[XI(i, j)] = arg;
            c->XR[XI(i, j)] = xr;
            for (int v = 0; v < 2; ++v) {
                float ar = ATT(i, j, v) + DEC(i, RIGHT, v, GO);
                AT(IR, i, j, v) = xr + (REAL)(sizeof(REAL) == 4 ? ar : (double)ATT(i, j, v) + (double)DEC(i, RIGHT, v, GO));
            }
        }
        for (int i = 0; i + w < N; ++i) {
            int j = i + w;
            for (int v = 0; v < 2; ++v) {
                /* step 3, dmv.py:58-59 */
                for (int r = i; r < j; ++r) scratch[r - i] = AT(CL, r, i, NOCHILD) + AT(IL, j, r, v);
                AT(CL, j, i, v) = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
                if (bpCL) bpCL[XI(j, i) * 2 + v] = arg;
                /* step 4, dmv.py:61-62 */
                for (int r = i + 1; r <= j; ++r) scratch[r - i - 1] = AT(IR, i, r, v) + AT(CR, r, j, NOCHILD);
                AT(CR, i, j, v) = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
                if (bpCR) bpCR[XI(i, j) * 2 + v] = arg;
            }
        }
        /* single-root mask, dmv.py:63 (value = class attribute `zero`) */
        if (w != len) {
            AT(CR, 0, w, 0) = mask_zero;
            AT(CR, 0, w, 1) = mask_zero;
        }
    }
}

/*
 * Log semiring: partition Z[b] and (optionally) d(sum_b gZ[b] Z[b]) / d(dec, attach)
 * = expected counts / arc marginals (helpers.py:150-154 done by hand).
 * dec    [B][N][2][2][2]  (pos, dir, val, decision)      merged (with ROOT at 0)
 * attach [B][N][N][2]     (head, child, val)             merged
 * lengths[B]              words without ROOT
 * trim   0: sweep the padded chart literally as the reference does
 *        1: sweep only positions 0..len (must give the same answer)
 */
int FN(vlgae_oracle_dmv_log)(const float *dec_all, const float *attach_all, const int64_t *lengths, int B, int N,
                             float mask_zero, int trim, const REAL *gZ, REAL *Z, REAL *gdec_all, REAL *gattach_all) {
    chart_t c, g;
    if (!chart_alloc(&c, N) || !chart_alloc(&g, N)) return 1;
    REAL *scratch = (REAL *)malloc(sizeof(REAL) * (size_t)(N + 1));
    for (int b = 0; b < B; ++b) {
        const float *dec = dec_all + (size_t)b * N * 8;
        const float *attach = attach_all + (size_t)b * N * N * 2;
        int len = (int)lengths[b];
        if (len < 0 || len > N - 1) { chart_free(&c); chart_free(&g); free(scratch); return 2; }
        int Nb = trim ? len + 1 : N;
        /* run on a compact [Nb][Nb] chart; inputs keep stride N */
        forward_one(dec, attach, N, Nb, len, (REAL)mask_zero, 0, &c, 0, 0, 0, 0, scratch);
        const int CS = Nb;
        Z[b] = AT(c.CR, 0, len, NOCHILD); /* dmv.py:65 */
        if (!gdec_all && !gattach_all) continue;
        REAL *gdec = gdec_all ? gdec_all + (size_t)b * N * 8 : 0;
        REAL *gatt = gattach_all ? gattach_all + (size_t)b * N * N * 2 : 0;
        if (gdec) memset(gdec, 0, sizeof(REAL) * (size_t)N * 8);
        if (gatt) memset(gatt, 0, sizeof(REAL) * (size_t)N * N * 2);
        {
            const int Ns = N;
            size_t n2 = (size_t)Nb * Nb;
            memset(g.CL, 0, sizeof(REAL) * n2 * 2); memset(g.CR, 0, sizeof(REAL) * n2 * 2);
            memset(g.IL, 0, sizeof(REAL) * n2 * 2); memset(g.IR, 0, sizeof(REAL) * n2 * 2);
            AT(g.CR, 0, len, NOCHILD) = gZ ? gZ[b] : (REAL)1;
            for (int w = Nb - 1; w >= 1; --w) {
                /* the mask overwrote CR[0][w] when w != len: no gradient passes through it */
                for (int i = 0; i + w < Nb; ++i) {
                    int j = i + w;
                    for (int v = 0; v < 2; ++v) {
                        /* step 4 transposed */
                        REAL gg = AT(g.CR, i, j, v);
                        if (i == 0 && w != len) gg = 0;
                        if (gg != 0) {
                            /* the forward value before masking: recompute the LSE for masked cells is
                               unnecessary because gg == 0 there */
                            REAL out = AT(c.CR, i, j, v);
                            for (int r = i + 1; r <= j; ++r) {
                                REAL p = gg * real_exp(AT(c.IR, i, r, v) + AT(c.CR, r, j, NOCHILD) - out);
                                AT(g.IR, i, r, v) += p;
                                AT(g.CR, r, j, NOCHILD) += p;
                            }
                        }
                        /* step 3 transposed */
                        gg = AT(g.CL, j, i, v);
                        if (gg != 0) {
                            REAL out = AT(c.CL, j, i, v);
                            for (int r = i; r < j; ++r) {
                                REAL p = gg * real_exp(AT(c.CL, r, i, NOCHILD) + AT(c.IL, j, r, v) - out);
                                AT(g.CL, r, i, NOCHILD) += p;
                                AT(g.IL, j, r, v) += p;
                            }
                        }
                    }
                }
                for (int i = 0; i + w < Nb; ++i) {
                    int j = i + w;
                    /* step 2 transposed */
                    REAL gx = 0;
                    for (int v = 0; v < 2; ++v) {
                        REAL gi = AT(g.IR, i, j, v);
                        gx += gi;
                        if (gatt) gatt[((size_t)i * Ns + j) * 2 + v] += gi;
                        if (gdec) gdec[((i * 2 + RIGHT) * 2 + v) * 2 + GO] += gi;
                    }
                    if (gx != 0) {
                        REAL out = c.XR[XI(i, j)];
                        for (int r = i; r < j; ++r) {
                            REAL p = gx * real_exp(AT(c.CR, i, r, HASCHILD) + AT(c.CL, j, r + 1, NOCHILD) - out);
                            AT(g.CR, i, r, HASCHILD) += p;
                            AT(g.CL, j, r + 1, NOCHILD) += p;
                        }
                    }
                    /* step 1 transposed */
                    gx = 0;
                    for (int v = 0; v < 2; ++v) {
                        REAL gi = AT(g.IL, j, i, v);
                        gx += gi;
                        if (gatt) gatt[((size_t)j * Ns + i) * 2 + v] += gi;
                        if (gdec) gdec[((j * 2 + LEFT) * 2 + v) * 2 + GO] += gi;
                    }
                    if (gx != 0) {
                        REAL out = c.XL[XI(j, i)];
                        for (int r = i; r < j; ++r) {
                            REAL p = gx * real_exp(AT(c.CR, i, r, NOCHILD) + AT(c.CL, j, r + 1, HASCHILD) - out);
                            AT(g.CR, i, r, NOCHILD) += p;
                            AT(g.CL, j, r + 1, HASCHILD) += p;
                        }
                    }
                }
            }
            if (gdec)
                for (int i = 0; i < Nb; ++i)
                    for (int v = 0; v < 2; ++v) {
                        gdec[((i * 2 + LEFT) * 2 + v) * 2 + STOP] = AT(g.CL, i, i, v);
                        gdec[((i * 2 + RIGHT) * 2 + v) * 2 + STOP] = AT(g.CR, i, i, v);
                    }
        }
    }
    chart_free(&c); chart_free(&g); free(scratch);
    return 0;
}

/*
 * Max semiring: best score, head of every word, dense arc indicator and the
 * decision-count "gradient" (what autograd through torch.max produces).
 * heads   [B][N]      heads[b][c] = h for 1 <= c <= len, 0 elsewhere
 * arcs    [B][N][N][2] 0/1 indicator (head, child, valence)   (may be NULL)
 * gdec    [B][N][2][2][2] decision counts of the best tree     (may be NULL)
 */
int FN(vlgae_oracle_dmv_viterbi)(const float *dec_all, const float *attach_all, const int64_t *lengths, int B, int N,
                                 float mask_zero, int trim, REAL *best, int64_t *heads_all, REAL *arcs_all,
                                 REAL *gdec_all) {
    chart_t c;
    if (!chart_alloc(&c, N)) return 1;
    size_t n2 = (size_t)N * N;
    int *bpXL = (int *)malloc(sizeof(int) * n2), *bpXR = (int *)malloc(sizeof(int) * n2);
    int *bpCL = (int *)malloc(sizeof(int) * n2 * 2), *bpCR = (int *)malloc(sizeof(int) * n2 * 2);
    int *stack = (int *)malloc(sizeof(int) * 4 * (size_t)(4 * N + 8));
    REAL *scratch = (REAL *)malloc(sizeof(REAL) * (size_t)(N + 1));
    for (int b = 0; b < B; ++b) {
        const float *dec = dec_all + (size_t)b * N * 8;
        const float *attach = attach_all + (size_t)b * N * N * 2;
        int len = (int)lengths[b];
        if (len < 0 || len > N - 1) return 2;
        int Nb = trim ? len + 1 : N;
        forward_one(dec, attach, N, Nb, len, (REAL)mask_zero, 1, &c, bpXL, bpXR, bpCL, bpCR, scratch);
        int64_t *heads = heads_all ? heads_all + (size_t)b * N : 0;
        REAL *arcs = arcs_all ? arcs_all + (size_t)b * N * N * 2 : 0;
        REAL *gdec = gdec_all ? gdec_all + (size_t)b * N * 8 : 0;
        if (heads) memset(heads, 0, sizeof(int64_t) * (size_t)N);
        if (arcs) memset(arcs, 0, sizeof(REAL) * n2 * 2);
        if (gdec) memset(gdec, 0, sizeof(REAL) * (size_t)N * 8);
        const int Ns = N;
        const int CS = Nb;
        best[b] = AT(c.CR, 0, len, NOCHILD);
        /* follow the first-max back-pointers from CR[0][len][NOCHILD]
           item kinds: 0 = CR(i,j,v) 1 = CL(j,i,v) 2 = IR(i,j,v) 3 = IL(j,i,v); fields (kind, lo, hi, v) */
        int sp = 0;
        stack[sp++] = 0; stack[sp++] = 0; stack[sp++] = len; stack[sp++] = NOCHILD;
        while (sp) {
            int v = stack[--sp], hi = stack[--sp], lo = stack[--sp], kind = stack[--sp];
            int i = lo, j = hi;
            if (kind == 0) {
                if (i == j) { if (gdec) gdec[((i * 2 + RIGHT) * 2 + v) * 2 + STOP] += 1; continue; }
                int r = i + 1 + bpCR[XI(i, j) * 2 + v];
                stack[sp++] = 2; stack[sp++] = i; stack[sp++] = r; stack[sp++] = v;
                stack[sp++] = 0; stack[sp++] = r; stack[sp++] = j; stack[sp++] = NOCHILD;
            } else if (kind == 1) {
                if (i == j) { if (gdec) gdec[((i * 2 + LEFT) * 2 + v) * 2 + STOP] += 1; continue; }
                int r = i + bpCL[XI(j, i) * 2 + v];
                stack[sp++] = 1; stack[sp++] = i; stack[sp++] = r; stack[sp++] = NOCHILD;
                stack[sp++] = 3; stack[sp++] = r; stack[sp++] = j; stack[sp++] = v;
            } else if (kind == 2) { /* arc i -> j */
                if (heads) heads[j] = i;
                if (arcs) arcs[((size_t)i * Ns + j) * 2 + v] = 1;
                if (gdec) gdec[((i * 2 + RIGHT) * 2 + v) * 2 + GO] += 1;
                int r = i + bpXR[XI(i, j)];
                stack[sp++] = 0; stack[sp++] = i; stack[sp++] = r; stack[sp++] = HASCHILD;
                stack[sp++] = 1; stack[sp++] = r + 1; stack[sp++] = j; stack[sp++] = NOCHILD;
            } else { /* arc j -> i */
                if (heads) heads[i] = j;
                if (arcs) arcs[((size_t)j * Ns + i) * 2 + v] = 1;
                if (gdec) gdec[((j * 2 + LEFT) * 2 + v) * 2 + GO] += 1;
                int r = i + bpXL[XI(j, i)];
                stack[sp++] = 0; stack[sp++] = i; stack[sp++] = r; stack[sp++] = NOCHILD;
                stack[sp++] = 1; stack[sp++] = r + 1; stack[sp++] = j; stack[sp++] = HASCHILD;
            }
        }
    }
    chart_free(&c);
    free(bpXL); free(bpXR); free(bpCL); free(bpCR); free(stack); free(scratch);
    return 0;
}

/*
 * Arc-factored projective CRF used for MBR decoding (deptree.py:25-76).
 * arc [B][N][N] (head, child); positions beyond len are masked with `fill`
 * (deptree.py:159-161), C[i][i] = 0, single-root mask value `mask_zero`.
 * semiring 0: log -> out[b] = log Z, marg = arc marginals
 * semiring 1: max -> out[b] = best,  marg = 0/1 indicator, heads filled
 */
int FN(vlgae_oracle_deptree)(const float *arc_all, const int64_t *lengths, int B, int N, float fill, float mask_zero,
                             int semiring, REAL *out, REAL *marg_all, int64_t *heads_all) {
    size_t n2 = (size_t)N * N;
    REAL *C = (REAL *)malloc(sizeof(REAL) * n2), *I = (REAL *)malloc(sizeof(REAL) * n2);
    REAL *X = (REAL *)malloc(sizeof(REAL) * n2), *A = (REAL *)malloc(sizeof(REAL) * n2);
    REAL *gC = (REAL *)malloc(sizeof(REAL) * n2), *gI = (REAL *)malloc(sizeof(REAL) * n2);
    int *bpX = (int *)malloc(sizeof(int) * n2), *bpC = (int *)malloc(sizeof(int) * n2);
    REAL *scratch = (REAL *)malloc(sizeof(REAL) * (size_t)(N + 1));
#define M(Q, a, b) Q[(size_t)(a) * N + (b)]
    for (int b = 0; b < B; ++b) {
        int len = (int)lengths[b];
        if (len < 0 || len > N - 1) return 2;
        const float *arc = arc_all + (size_t)b * n2;
        for (int h = 0; h < N; ++h)
            for (int cc = 0; cc < N; ++cc) M(A, h, cc) = (h > len || cc > len) ? (REAL)fill : (REAL)M(arc, h, cc);
        for (size_t t = 0; t < n2; ++t) { C[t] = fill; I[t] = fill; }
        for (int i = 0; i < N; ++i) M(C, i, i) = 0;
        int arg = 0;
        for (int w = 1; w < N; ++w) {
            for (int i = 0; i + w < N; ++i) {
                int j = i + w;
                for (int r = i; r < j; ++r) scratch[r - i] = M(C, i, r) + M(C, j, r + 1);
                REAL x = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
                M(X, i, j) = x; M(bpX, i, j) = arg;
                M(I, j, i) = x + M(A, j, i);
                M(I, i, j) = x + M(A, i, j);
            }
            for (int i = 0; i + w < N; ++i) {
                int j = i + w;
                for (int r = i; r < j; ++r) scratch[r - i] = M(C, r, i) + M(I, j, r);
                M(C, j, i) = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
                M(bpC, j, i) = arg;
                for (int r = i + 1; r <= j; ++r) scratch[r - i - 1] = M(I, i, r) + M(C, r, j);
                M(C, i, j) = semiring ? first_max(scratch, w, &arg) : lse(scratch, w);
                M(bpC, i, j) = arg;
            }
            if (w != len) M(C, 0, w) = mask_zero;
        }
        out[b] = M(C, 0, len);
        REAL *marg = marg_all ? marg_all + (size_t)b * n2 : 0;
        int64_t *heads = heads_all ? heads_all + (size_t)b * N : 0;
        if (heads) memset(heads, 0, sizeof(int64_t) * (size_t)N);
        if (!marg && !heads) continue;
        for (size_t t = 0; t < n2; ++t) { gC[t] = 0; gI[t] = 0; }
        M(gC, 0, len) = 1;
        for (int w = N - 1; w >= 1; --w) {
            for (int i = 0; i + w < N; ++i) {
                int j = i + w;
                REAL gg = M(gC, i, j);
                if (i == 0 && w != len) gg = 0;
                if (gg != 0) {
                    for (int r = i + 1; r <= j; ++r) {
                        REAL p = semiring ? (REAL)(r - i - 1 == M(bpC, i, j)) * gg
                                          : gg * real_exp(M(I, i, r) + M(C, r, j) - M(C, i, j));
                        M(gI, i, r) += p; M(gC, r, j) += p;
                    }
                }
                gg = M(gC, j, i);
                if (gg != 0) {
                    for (int r = i; r < j; ++r) {
                        REAL p = semiring ? (REAL)(r - i == M(bpC, j, i)) * gg
                                          : gg * real_exp(M(C, r, i) + M(I, j, r) - M(C, j, i));
                        M(gC, r, i) += p; M(gI, j, r) += p;
                    }
                }
            }
            for (int i = 0; i + w < N; ++i) {
                int j = i + w;
                REAL gx = M(gI, i, j) + M(gI, j, i);
                if (heads && semiring) {
                    if (M(gI, i, j) != 0) heads[j] = i;
                    if (M(gI, j, i) != 0) heads[i] = j;
                }
                if (gx != 0) {
                    for (int r = i; r < j; ++r) {
                        REAL p = semiring ? (REAL)(r - i == M(bpX, i, j)) * gx
                                          : gx * real_exp(M(C, i, r) + M(C, j, r + 1) - M(X, i, j));
                        M(gC, i, r) += p; M(gC, j, r + 1) += p;
                    }
                }
            }
        }
        if (marg) {
            for (size_t t = 0; t < n2; ++t) marg[t] = gI[t];
            for (int i = 0; i < N; ++i) M(marg, i, i) = 0;
        }
    }
#undef M
    free(C); free(I); free(X); free(A); free(gC); free(gI); free(bpX); free(bpC); free(scratch);
    return 0;
}
