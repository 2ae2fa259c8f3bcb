// align_bwd_kernels.cu -- backward of the alignment logits on the sm_100a tensor cores.
//
// The reference's attmap = einsum("avd,bqd->baqv") + two masked_fill_ (/root/reference/src/model/joint.py:413-418) is an
// autograd node; with g = d loss / d attmap its backward is the two transposed contractions
//     d txt[b,q,:] = m_q[b,q] * sum_{a,v} g[b,a,q,v] * m_v[a,v] * vis[a,v,:]          (KIND 0)
//     d vis[a,v,:] = m_v[a,v] * sum_{b,q} g[b,a,q,v] * m_q[b,q] * txt[b,q,:]          (KIND 1)
// (a masked entry was overwritten, so it passes no gradient).  Both stream the 7.4 GB gradient once.
//
// One kernel template serves both.  A step handles one [nq queries x 128 factors] tile of g:
//   * 16 converter warps read it with plain coalesced loads (its rows are 4-byte aligned only: V is odd), apply the mask
//     of the OTHER operand's side, split every value into bf16 hi + lo and store the 128-byte-swizzled K-major image
//     [part][64-factor block][query row][64 factors];
//   * warp 0 brings the matching packed operand tile (vis tile of (a, v-tile) / caption tile of b) with bulk TMA -- the
//     same images align_pack_kernel builds for the forward pass;
//   * warp 1 issues hi*hi + lo*hi + hi*lo with the accumulator resident in tensor memory for the whole work item:
//       KIND 0: D[d][q] += vis_tile^T x g_tile^T   A = vis tile read MN-major (M = d contiguous, K = factor rows),
//                                                  B = g image K-major   (N = query rows, K = factors)
//       KIND 1: D[v][d] += g_tile^T  x txt_tile    A = g image read MN-major (M = factors contiguous, K = query rows),
//                                                  B = caption tile read MN-major (N = d contiguous, K = query rows)
//     The K-major images of the forward pass ARE the MN-major images of the transposed products: a row of 64 contiguous
//     bf16 is one 128-byte line of the swizzle atom either way; only the descriptor's major bit and strides change.
//   * after the last step of an item 4 warps read the accumulator, apply the output-side mask and store.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "align_kernels.cuh"
#include "dmv_kernels.cuh"

namespace vlgae {
namespace {

constexpr int TILE = 128;
constexpr int CHUNK = TILE * 128;      // one (part, 64-column block) chunk of a packed operand tile: 128 rows x 128 B
constexpr int kConvWarps = 16;
constexpr int kBwdThreads = 64 + 32 * kConvWarps;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BW_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra BW_DONE_%=;\n\t"
        "bra BW_WAIT_%=;\n\t"
        "BW_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)) : "memory");
}
// both operands from shared memory; `acc` = 0 overwrites the accumulator (first MMA of a work item)
__device__ __forceinline__ void tc_mma_ss_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptors, 128-byte swizzle, sm_100 version bits.
//   K-major : rows of 64 K-elements; 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: rows of 64 MN-elements; the K index walks the rows (8-row groups 1024 B apart, SBO), the next 64 MN-elements
//             are `lbo_bytes` away (LBO) -- canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t saddr, uint32_t lbo_bytes) {
    const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// D = f32, A = B = bf16; bit 15 / 16 = A / B is MN-major
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct BwdSmem {
    uint64_t op_full[2], g_full[2], empty[2], acc_full, acc_empty;
    uint32_t tmem_base;
};

struct BwdArgs {
    const float *g;                // [B][A][Q][ldg] upstream gradient
    const uint8_t *op_packed;      // KIND 0: vis tiles [A][VT][2*KB][128 x 128 B]; KIND 1: caption tiles [B][QT][2*KB][128 x 128 B]
    const uint8_t *vis_mask;       // [A][V]
    const uint8_t *txt_mask;       // [B][Q]
    float *out;                    // KIND 0: d txt [B][Q][D]; KIND 1: d vis [A][V][D]
    int ldg, A, V, B, Q, D, KB, VT, QT, nq, split;
    int stages;                    // 2, or 1 when two stages do not fit (Q tiles of 128 queries with D = 128)
    int ksplit;                    // work units per item: each reduces a contiguous range of the steps and ADDS its partial
    long long *prof;               // debug: per CTA clocks of the MMA warp (total, waiting for the operand tile, for the g image)
};

// KIND 0: item = (b, qt), steps = (a, vt);  KIND 1: item = (a, vt), steps = (b, qt)
template <int KIND>
__global__ void __launch_bounds__(kBwdThreads, 1) align_bwd_kernel(BwdArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = p.KB, nq = p.nq;
    const uint32_t chunk_g = (uint32_t)nq * 128u;                       // one (part, 64-factor block) chunk of the g image
    const uint32_t g_bytes = 4u * chunk_g;                               // [hi, lo] x [factors 0..63, 64..127]
    const uint32_t chunk_op = KIND == 0 ? (uint32_t)CHUNK : chunk_g;     // rows kept of an operand chunk: 128 factors / nq queries
    const uint32_t op_bytes = 2u * KB * chunk_op;
    // KIND 0 reads the operand MN-major with M = 128 = two 64-wide blocks; with D <= 64 (KB = 1) the second block of the
    // `lo` part lies one chunk behind the tile (its rows land in accumulator lanes >= 64, never stored): keep it in range
    const uint32_t op_alloc = KIND == 0 ? 4u * chunk_op : op_bytes;
    const uint32_t stage_bytes = op_alloc + g_bytes;
    const uint32_t S = (uint32_t)p.stages;
    BwdSmem *sb = reinterpret_cast<BwdSmem *>(smem + (size_t)S * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&sb->op_full[s], 1); mbar_init(&sb->g_full[s], kConvWarps); mbar_init(&sb->empty[s], 1); }
        mbar_init(&sb->acc_full, 1);
        mbar_init(&sb->acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sb->tmem_base;

    const int n_items = KIND == 0 ? p.B * p.QT : p.A * p.VT;
    const int n_steps = KIND == 0 ? p.A * p.VT : p.B * p.QT;
    const int NOUT = KIND == 0 ? nq : 64 * KB;   // accumulator columns (MMA N)
    // A work unit = (item, k-slice): with KS > 1 an item's steps are cut into KS contiguous ranges whose partial results
    // are added with fp32 atomics (the output is zeroed by the launcher).  That evens out the waves (128 captions on 148
    // SMs leave 14 % of the machine idle) and shortens the chains of truncating tensor-core accumulations 15-fold.
    const int KS = p.ksplit, n_units = n_items * KS;
    auto unit = [&](int u, int &item, int &s0, int &s1) {
        item = u / KS;
        const int ks = u - item * KS;
        s0 = (int)(((long long)ks * n_steps) / KS);
        s1 = (int)(((long long)(ks + 1) * n_steps) / KS);
    };

    if (warp == 0) {
        // ===================== TMA producer: the packed operand tile of every step =====================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int item, s0, s1;
                unit(u, item, s0, s1);
                for (int step = s0; step < s1; ++step) {
                    mbar_wait(&sb->empty[s], ph ^ 1);
                    mbar_expect_tx(&sb->op_full[s], op_bytes);
                    const uint8_t *src = p.op_packed + (size_t)step * (size_t)(2 * KB * CHUNK);
                    uint8_t *dst = smem + (size_t)s * stage_bytes;
                    for (int ch = 0; ch < 2 * KB; ++ch)
                        bulk_g2s(dst + (size_t)ch * chunk_op, src + (size_t)ch * CHUNK, chunk_op, &sb->op_full[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, elect.sync) =====================
        const uint32_t idesc = KIND == 0 ? idesc_bf16(TILE, nq, true, false) : idesc_bf16(TILE, 64 * KB, true, true);
        uint32_t s = 0, ph = 0, it = 0;
        long long t_op = 0, t_g = 0;
        const long long t_begin = clock64();
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
            int item, s0, s1;
            unit(u, item, s0, s1);
            mbar_wait(&sb->acc_empty, (it & 1) ^ 1);
            tc_fence_after();
            for (int step = s0; step < s1; ++step) {
                const long long t0 = clock64();
                mbar_wait(&sb->op_full[s], ph);
                const long long t1 = clock64();
                mbar_wait(&sb->g_full[s], ph);
                t_op += t1 - t0;
                t_g += clock64() - t1;
                tc_fence_after();
                const uint32_t op = smem_u32(smem + (size_t)s * stage_bytes), gi = op + op_alloc;
                // hi*hi + lo*hi + hi*lo   (operand tile parts: chunk part * KB + block; g image parts: chunk part * 2 + block)
#pragma unroll 1
                for (int term = 0; term < 3; ++term) {
                    if (term > 0 && p.split == 1) break;
                    const int p_op = term == 1 ? 1 : 0, p_g = term == 2 ? 1 : 0;
                    if (KIND == 0) {
                        // A = vis tile, MN-major: M = d (KB blocks of 64, LBO = one chunk), K = 128 factor rows, 16 per MMA
                        // B = g image, K-major: N = query rows, K = factors: block k / 4, 32 B per MMA inside the atom
                        const uint32_t a0 = op + (uint32_t)(p_op * KB) * chunk_op, b0 = gi + (uint32_t)(p_g * 2) * chunk_g;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t ad = desc_mn_major(a0 + (uint32_t)k * 2048u, chunk_op);
                            const uint64_t bd = desc_k_major(b0 + (uint32_t)(k >> 2) * chunk_g + (uint32_t)(k & 3) * 32u);
                            tc_mma_ss_elect(tmem_base, ad, bd, idesc, ((step - s0) | term | k) != 0);
                        }
                    } else {
                        // A = g image, MN-major: M = factors (2 blocks of 64, LBO = one chunk), K = query rows, 16 per MMA
                        // B = caption tile, MN-major: N = d (KB blocks of 64), K = query rows
                        const uint32_t a0 = gi + (uint32_t)(p_g * 2) * chunk_g, b0 = op + (uint32_t)(p_op * KB) * chunk_op;
                        for (int k = 0; k < nq / 16; ++k) {
                            const uint64_t ad = desc_mn_major(a0 + (uint32_t)k * 2048u, chunk_g);
                            const uint64_t bd = desc_mn_major(b0 + (uint32_t)k * 2048u, chunk_op);
                            tc_mma_ss_elect(tmem_base, ad, bd, idesc, ((step - s0) | term | k) != 0);
                        }
                    }
                }
                tc_commit_elect(&sb->empty[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
            tc_commit_elect(&sb->acc_full);
        }
        if (p.prof && lane == 0) {
            long long *o = p.prof + (size_t)blockIdx.x * 8;
            o[0] = clock64() - t_begin; o[1] = t_op; o[2] = t_g;
        }
    } else {
        // ===================== converters (all 16 warps) + epilogue (the first 4) =====================
        const int cw = warp - 2;
        uint32_t s = 0, ph = 0, it = 0;
        constexpr int kRows = TILE / kConvWarps;  // rows per warp (nq <= 128)
        // Loads of one step: this warp's rows q = cw + 16 r of the [queries x 128 factors] tile; a lane owns the factor pairs
        // (64 j + 2 lane, + 1), j = 0, 1, so that hi and lo leave as ONE packed bf16x2 store each.  The other side's mask is
        // folded in here (KIND 0: m_v per column, KIND 1: m_q per row); out-of-range elements are zero.  All loads of a step
        // are issued together and one step ahead of their conversion: the tile comes cold from HBM.
        auto load_step = [&](int item, int step, float (&x)[kRows][4]) {
            int a, vt, b, qt;
            if (KIND == 0) { b = item / p.QT; qt = item - b * p.QT; a = step / p.VT; vt = step - a * p.VT; }
            else { a = item / p.VT; vt = item - a * p.VT; b = step / p.QT; qt = step - b * p.QT; }
            const int q_lim = min(TILE, p.Q - qt * TILE);
            bool colk[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int v = vt * TILE + 64 * (c >> 1) + 2 * lane + (c & 1);
                colk[c] = v < p.V && (KIND == 1 || p.vis_mask[(size_t)a * p.V + v] != 0);
            }
            const float *gsrc = p.g + (((size_t)b * p.A + a) * p.Q + (size_t)qt * TILE) * p.ldg + (size_t)vt * TILE + 2 * lane;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const int q = cw + r * kConvWarps;
                const bool rowk = q < q_lim && (KIND == 0 || p.txt_mask[(size_t)b * p.Q + qt * TILE + q] != 0);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    x[r][c] = (rowk && colk[c]) ? __ldcs(gsrc + (size_t)q * p.ldg + 64 * (c >> 1) + (c & 1)) : 0.f;
            }
        };
        // byte offset of this lane's pair inside a 128-byte row of the swizzled image: 16-byte unit (lane >> 2) ^ (q & 7),
        // and q & 7 == cw & 7 for every row of this warp (rows are 16 apart)
        const uint32_t lane_off = (uint32_t)((((lane >> 2) ^ (cw & 7)) << 4) + (lane & 3) * 4 + cw * 128);
        float x[kRows][4], xn[kRows][4];
        if ((int)blockIdx.x < n_units) {
            int item, s0, s1;
            unit(blockIdx.x, item, s0, s1);
            load_step(item, s0, x);
        }
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
            int item, s0, s1;
            unit(u, item, s0, s1);
            for (int step = s0; step < s1; ++step) {
                // the next step's tile (possibly the next unit's first) is requested before this one is converted
                if (step + 1 < s1) load_step(item, step + 1, xn);
                else if (u + (int)gridDim.x < n_units) {
                    int item2, t0, t1;
                    unit(u + gridDim.x, item2, t0, t1);
                    load_step(item2, t0, xn);
                }
                mbar_wait(&sb->empty[s], ph ^ 1);
                uint8_t *gi = smem + (size_t)s * stage_bytes + op_alloc + lane_off;
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if (cw + r * kConvWarps < nq) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(x[r][2 * j], x[r][2 * j + 1]);
                            const float2 hf = __bfloat1622float2(h);
                            const __nv_bfloat162 l = __floats2bfloat162_rn(x[r][2 * j] - hf.x, x[r][2 * j + 1] - hf.y);
                            uint8_t *dst = gi + (uint32_t)j * chunk_g + (uint32_t)(r * kConvWarps) * 128u;
                            *reinterpret_cast<__nv_bfloat162 *>(dst) = h;
                            *reinterpret_cast<__nv_bfloat162 *>(dst + 2u * chunk_g) = l;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) x[r][c] = xn[r][c];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&sb->g_full[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
            if (cw < 4) {
                // epilogue: accumulator lane = warp quadrant * 32 + lane
                mbar_wait(&sb->acc_full, it & 1);
                tc_fence_after();
                const int quad = warp & 3, row = quad * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
                if (KIND == 0) {
                    const int b = item / p.QT, qt = item - b * p.QT, q_lim = min(TILE, p.Q - qt * TILE);
                    for (int c0 = 0; c0 < NOUT; c0 += 16) {  // lane = d, columns = queries
                        uint32_t r[16];
                        tc_ld16(taddr + (uint32_t)c0, r);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int q = c0 + j;
                            if (q < q_lim && row < p.D) {
                                const size_t qq = (size_t)b * p.Q + qt * TILE + q;
                                const float val = p.txt_mask[qq] ? __uint_as_float(r[j]) : 0.f;
                                if (KS > 1) atomicAdd(p.out + qq * p.D + row, val);
                                else p.out[qq * p.D + row] = val;
                            }
                        }
                    }
                } else {
                    const int a = item / p.VT, vt = item - a * p.VT, v = vt * TILE + row;
                    const bool keep = v < p.V && p.vis_mask[(size_t)a * p.V + v];
                    for (int c0 = 0; c0 < NOUT; c0 += 16) {  // lane = factor, columns = d
                        uint32_t r[16];
                        tc_ld16(taddr + (uint32_t)c0, r);
                        if (v < p.V) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < p.D) {
                                    const float val = keep ? __uint_as_float(r[j]) : 0.f;
                                    if (KS > 1) atomicAdd(p.out + ((size_t)a * p.V + v) * p.D + c0 + j, val);
                                    else p.out[((size_t)a * p.V + v) * p.D + c0 + j] = val;
                                }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sb->acc_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

}  // namespace

// packs both operands (workspace layout of launch_align) and runs the requested backward kernels
cudaError_t launch_align_backward(const float *g, int ldg, const float *vis, const uint8_t *vis_mask, const float *txt,
                                  const uint8_t *txt_mask, int A, int V, int B, int Q, int D, int split, float *grad_vis,
                                  float *grad_txt, void *workspace, cudaStream_t st) {
    const AlignPlan pl = align_plan(A, V, B, Q, D);
    uint8_t *ws = reinterpret_cast<uint8_t *>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    uint8_t *vis_packed = ws, *txt_packed = vis_packed + pl.vis_packed_bytes;
    cudaError_t e = align_pack_operands(vis, txt, txt_mask, A, V, B, Q, D, workspace, st);
    if (e != cudaSuccess) return e;
    int dev = 0, sm = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    BwdArgs a{};
    a.g = g; a.ldg = ldg; a.vis_mask = vis_mask; a.txt_mask = txt_mask;
    a.A = A; a.V = V; a.B = B; a.Q = Q; a.D = D; a.KB = pl.KB; a.VT = pl.VT; a.QT = pl.QT; a.nq = pl.nq;
    a.split = split == 1 ? 1 : 3;
    a.prof = dmv_profile_buffer();
    const size_t g_bytes = (size_t)4 * pl.nq * 128;
    // k-slices per item: the count (<= 16, <= steps) that fills the last wave of persistent CTAs best
    auto pick_ksplit = [&](int items, int steps) {
        int best = 1;
        double best_eff = 0.0;
        for (int ks = 1; ks <= 16 && ks <= steps; ++ks) {
            const long long units = (long long)items * ks, waves = (units + sm - 1) / sm;
            const double eff = (double)units / (double)(waves * sm);
            if (eff > best_eff + 0.02) { best_eff = eff; best = ks; }
        }
        return best;
    };
    auto run = [&](auto kern, size_t stage, int items, int steps, float *out, size_t out_elems) -> cudaError_t {
        a.ksplit = pick_ksplit(items, steps);
        if (a.ksplit > 1) {
            cudaError_t err = cudaMemsetAsync(out, 0, out_elems * sizeof(float), st);
            if (err != cudaSuccess) return err;
        }
        items *= a.ksplit;
        const size_t fixed = sizeof(BwdSmem) + 64;
        a.stages = 2 * stage + fixed <= (size_t)smem_max ? 2 : 1;
        const size_t smem = a.stages * stage + fixed;
        if (smem > (size_t)smem_max) return cudaErrorInvalidValue;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<items < sm ? items : sm, kBwdThreads, smem, st>>>(a);
        return cudaGetLastError();
    };
    if (grad_txt) {
        a.op_packed = vis_packed; a.out = grad_txt;
        e = run(align_bwd_kernel<0>, (size_t)4 * CHUNK + g_bytes, B * pl.QT, A * pl.VT, grad_txt, (size_t)B * Q * D);
        if (e != cudaSuccess) return e;
    }
    if (grad_vis) {
        a.op_packed = txt_packed; a.out = grad_vis;
        e = run(align_bwd_kernel<1>, (size_t)2 * pl.KB * pl.nq * 128 + g_bytes, A * pl.VT, B * pl.QT, grad_vis, (size_t)A * V * D);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace vlgae
