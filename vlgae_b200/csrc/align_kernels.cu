// align_kernels.cu -- word/arc x scene-graph-factor alignment scores on the sm_100a tensor cores.
//
// Replaces  gather_logit_simple / gather_logit_reduced   /root/reference/src/model/joint.py:406-432
//   attmap[b, a, q, v] = < txt[b, q, :], vis[a, v, :] >,  then -INF where vis_mask[a, v] or txt_mask[b, q] is false.
// The reference runs a cuBLAS fp32 SGEMM plus two full passes of masked_fill_ over the 7.4 GB result; here the
// contraction, both masks and the store are one kernel, and the result is written exactly once.
//
// Mapping onto tcgen05:  D[v][q] = sum_d vis[a, v, d] * txt[b, q, d]
//   M = 128 factors v of one image a      (operand A, in TENSOR MEMORY: two tiles = 256 factors stay resident per CTA)
//   N = Q queries of one caption b, <= 128 (operand B, K-major, streamed through a ring of shared-memory slots)
//   K = D <= 128
// so that TMEM lane = v and TMEM column = q: the 32 threads of an epilogue warp hold 32 consecutive v of one q in
// the same register -- a conflict-free shared-memory row segment of the staged output tile [q][128 v].
//
// Precision: operands are split into bf16 hi + bf16 lo (x = hi + lo + O(2^-17 x)) and three MMAs accumulate
// hi*hi + lo*hi + hi*lo in fp32 TMEM -- fp32-class logits (|err| ~ 1e-5 |x||y| sqrt(D)) from the bf16 pipe.
//
// Data movement: a pre-pass (align_pack_kernel) writes both operands as bf16 tiles that are already the 128-byte
// swizzled shared-memory image tcgen05 expects, so the main kernel moves them with plain 1-D bulk TMA copies
// (cp.async.bulk ... mbarrier::complete_tx) -- no tensor maps. The result leaves through bulk TMA stores as well.
// Warp roles (18 warps, one persistent CTA per SM):
//   warp 0      TMA producer: image-tile chunks and caption tiles into the ring
//   warp 1      MMA issuer (converged warp, elect.sync): tcgen05.cp image tile -> TMEM, 24 TS-form MMAs per tile pair
//   warps 2-9   epilogue team 0, warps 10-17 epilogue team 1: alternate accumulators; TMEM -> registers -> vis mask ->
//               staged tile in shared memory -> one 512 B cp.async.bulk store per query row (masked queries copy a
//               constant row of -INF instead)
// What bounds it (cfg2 shape, measured): an SM's path to L2 takes ~19 B/clk of stores whichever way they are issued
// (5.3 TB/s for the chip at full clock, tools/probes/), and the tensor work holds the clock near 1.6 GHz.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "align_kernels.cuh"
#include "dmv_kernels.cuh"

namespace vlgae {
namespace {

constexpr int TILE_M = 128;          // factors per tile (TMEM lanes)
constexpr int CHUNK_A = TILE_M * 128;  // bytes of one (part, k-block) chunk of operand A: 128 rows x 64 bf16
constexpr int MAX_ACC = 2;           // TMEM accumulators behind the two image tiles, stride = nq rounded to 32
constexpr int kTeamWarps = 8;  // an epilogue team: 2 warps per TMEM lane quadrant, interleaved over 16-column chunks
constexpr int kTeams = 2;      // teams take alternate accumulators, so one stages / stores while the other unloads TMEM
constexpr int kEpiWarps = kTeamWarps;  // (per team)
constexpr int kThreads = 64 + 32 * kTeamWarps * kTeams;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
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
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Variants for a converged warp: every lane executes the call, one elected lane issues. Inside `if (lane == 0)` the
// compiler must treat the operands as per-thread values and wraps every tcgen05 instruction in an ELECT / R2UR / branch
// sequence (~8 instructions, ~70 clk per MMA -- more than the 48 clk the MMA itself takes); in converged code the
// descriptors stay in uniform registers.
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)) : "memory");
}
template <int ACC>
__device__ __forceinline__ void tc_mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void tc_cp_128x256b_elect(uint32_t dst_tmem, uint64_t sdesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.cp.cta_group::1.128x256b [%0], %1;\n\t"
        "}" ::"r"(dst_tmem), "l"(sdesc) : "memory");
}
// shared memory -> tensor memory: 128 rows x 256 bits (one UMMA_K slice of a K-major operand), lanes = rows

// 32 lanes x 16 columns of fp32: thread i of the warp gets lane (quadrant*32 + i), columns c0 .. c0+15
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (dense), sm_100 version
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFF) >> 4) | (1u << 16);  // start address | LBO = 1 (unused for swizzled K-major)
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B | version 1 | SWIZZLE_128B
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M x N
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// the same load without the wait: several can be in flight; tc_wait_ld ties the registers to the wait
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// shared -> global bulk copy (TMA), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(float *dst, const float *src, int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------
// pre-pass: fp32 [G][R][D] -> per (g, row tile of 128) bf16 hi / lo chunks in the swizzled smem image
//   chunk order inside a tile: [hi kb0][hi kb1]..[lo kb0][lo kb1]..; each chunk `rows_per_chunk` rows x 128 B
//   element (r, k): byte r*128 + (((k%64)/8) ^ (r%8))*16 + (k%8)*2 of chunk k/64
// also packs a bool mask [G][R] into bit words [G][tiles][4] (bit r%32 of word r/32 inside the tile)
// ---------------------------------------------------------------------------------------------
__global__ void align_pack_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, int G, int R, int D,
                                  int KB, int rows_per_chunk, size_t tile_stride_bytes, int tiles, uint8_t *out,
                                  uint32_t *maskbits, int tile_step) {
    // tile_step = first row of tile t is t * tile_step: 128, or 120 for the overlapping tiles of the sector-owning dense
    // layout (align_gemm_kernel, VSTEP)
    // one thread per (g, tile, row, 8-element group): writes one 16-byte unit of hi and of lo
    const int groups = KB * 8;
    const size_t total = (size_t)G * tiles * rows_per_chunk * groups;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int grp = (int)(t % groups);
        size_t u = t / groups;
        const int r = (int)(u % rows_per_chunk); u /= rows_per_chunk;
        const int tile = (int)(u % tiles);
        const int g = (int)(u / tiles);
        const int row = tile * tile_step + r;
        const int k0 = grp * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (row < R && k0 + e < D) ? x[((size_t)g * R + row) * D + k0 + e] : 0.f;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
            hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const int kb = grp >> 3, c = grp & 7;
        const size_t chunk_bytes = (size_t)rows_per_chunk * 128;
        uint8_t *tile_base = out + ((size_t)g * tiles + tile) * tile_stride_bytes;
        const size_t off = (size_t)r * 128 + (size_t)((c ^ (r & 7)) * 16);
        *reinterpret_cast<uint4 *>(tile_base + (size_t)kb * chunk_bytes + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(tile_base + (size_t)(KB + kb) * chunk_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (maskbits && grp == 0 && (r & 31) == 0) {
            uint32_t w = 0;
            for (int e = 0; e < 32; ++e) {
                const int rr = row + e;
                if (rr < R && mask[(size_t)g * R + rr]) w |= 1u << e;
            }
            maskbits[((size_t)g * tiles + tile) * 4 + (r >> 5)] = w;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct AlignSmem {
    uint64_t ring_full[8], ring_empty[8];
    uint64_t acc_full[MAX_ACC], acc_empty[MAX_ACC];
    uint32_t tmem_base;
};

// Shared memory:  [ring: S slots of slot_bytes][BULK: NB output tiles of nq x 128 fp32][AlignSmem]
//   A ring slot holds either one caption tile (2*KB chunks of nq rows x 128 B: hi k-blocks, lo k-blocks) or up to
//   slot_bytes / 16 KB chunks of an image tile on their way to tensor memory -- the image tiles travel through the same
//   ring as the captions, so no shared memory is reserved for them and the ring never drains between work items.
// Tensor memory (512 columns): [image tile 0: 64*KB columns][image tile 1][accumulators, stride = nq rounded to 32]
//   A CTA keeps TWO image tiles (256 factors) resident and runs every caption tile against both (TS-form MMA, operand A
//   from tensor memory): with one tile the kernel streams as many caption bytes L2 -> SM as it writes, and the L2
//   slices, which carry the reads, the writes and the write-back together, are what saturates.
// MODE 0: direct stores (rows not 16-byte aligned)   1: staged tile + bulk TMA stores   2: no [B,A,Q,V] output at all --
// the epilogue reduces every query row to its maximum over the factors (and the arg-max) and merges the v-tiles with a
// 64-bit atomicMax (gather_logit_reduced, joint.py:421-432: the 7.4 GB tensor is never materialised)
template <int KB, int MODE, int VSTEP = 0>
__global__ void __launch_bounds__(kThreads, 1) align_gemm_kernel(AlignArgs p) {
    constexpr bool BULK = MODE >= 1;      // a staged [query][128 factors] tile per epilogue team
    constexpr bool REDUCE = MODE >= 2;
    constexpr bool MAXQ = MODE == 3;      // REDUCE + the max over the queries (its own instantiation: the plain reduction
                                          // stays free of the extra registers, which spilled under the 96-register cap)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nq = p.nq, S = p.stages, NB = p.out_bufs;
    const uint32_t chunk_b = (uint32_t)nq * 128u;         // one (part, k-block) chunk of a caption tile
    const uint32_t txt_bytes = 2u * KB * chunk_b;           // a caption tile in a slot
    const uint32_t slot_bytes = p.slot_bytes;               // >= txt_bytes and >= one image chunk
    const int cps = (int)(slot_bytes / CHUNK_A);            // image chunks per slot
    constexpr int PAIR = 2;
    constexpr uint32_t A_COLS = 64u * KB;  // tensor-memory columns of one image tile (hi + lo, 32 per 64-wide k-block)
    uint8_t *s_ring = smem;
    float *s_out = reinterpret_cast<float *>(smem + (size_t)S * slot_bytes);
    // rows = min(nq, Q): the queries a tile can hold.  TS = row stride of a staged tile: 128 floats, or 132 in the SHIFTED
    // layout used when the rows of `out` are not 16-byte aligned (odd V, no padding): row r is staged k_r floats to the
    // right, k_r = word offset of its global destination modulo 4, so that shared and global addresses agree modulo 16 and
    // all but <= 3 leading / trailing floats of the row still leave through one bulk copy
    // (SHIFT is a template parameter: with a run-time stride the 16 staging stores per chunk lose their immediate offsets
    // and the aligned layout went from 1.76 to 3.2 ms)
    // VSTEP != 0 selects the shifted layout and is the distance between the first factors of consecutive tiles: 128
    // (disjoint tiles: every 512-byte row segment starts and ends inside a 32-byte sector that a neighbouring tile -- another
    // CTA, or this one later -- completes: 560 M L2 write-sector operations instead of 386 M, 3.1-3.3 ms) or 120: tiles
    // OVERLAP by 8 factors and tile t > 0 stores, of row r, the 120 floats [120 t + e_r, 120 (t + 1) + e_r), e_r < 8 chosen so
    // that the segment starts on a sector boundary of the global row -- every sector of a row except its first and its
    // last is then written whole, by one bulk copy.  1369 factors are 12 tiles either way it is counted in pairs (6).
    // (Owning whole 128-byte lines instead -- tiles 96 factors apart, the same code with U = 32 -- was measured too: 14 tiles =
    // 7 pairs, 2.43 ms against 2.25 ms; not instantiated.)
    constexpr bool SHIFT = VSTEP != 0;
    static_assert(!SHIFT || MODE == 1, "the shifted staging belongs to the materialising bulk-store mode");
    static_assert(VSTEP == 0 || VSTEP == TILE_M || VSTEP == TILE_M - 8 || VSTEP == TILE_M - 32, "tile step: 128, 120 or 96");
    constexpr int TS = SHIFT ? TILE_M + 4 : TILE_M;
    constexpr bool shifted = SHIFT;
    constexpr int VS = SHIFT ? VSTEP : TILE_M;  // factor step between tiles
    const size_t out_tile_floats = (size_t)p.out_rows * TS;
    float *s_neg = s_out + (size_t)NB * out_tile_floats;  // BULK: one row of -INF, the source of masked query rows
    // REDUCE: running (ordered max bits << 32 | ~arg-max) per (team, caption of the chunk, query); 0 = nothing seen yet
    unsigned long long *s_run = reinterpret_cast<unsigned long long *>(s_neg + TILE_M);
    AlignSmem *sb = reinterpret_cast<AlignSmem *>(smem + (size_t)S * slot_bytes + (BULK ? (size_t)NB * p.out_rows * TS * 4 + 512 : 0) +
                                                  (REDUCE ? (size_t)p.run_bytes : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&sb->ring_full[s], 1); mbar_init(&sb->ring_empty[s], 1); }
        for (int a = 0; a < MAX_ACC; ++a) { mbar_init(&sb->acc_full[a], 1); mbar_init(&sb->acc_empty[a], kTeamWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (BULK && threadIdx.x >= 64 && threadIdx.x < 64 + TILE_M) {
        s_neg[threadIdx.x - 64] = p.neg;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (REDUCE)
        for (int t = threadIdx.x; t < (int)(p.run_bytes / 8); t += kThreads) s_run[t] = 0ull;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // the swizzled tile images need 1024-byte alignment
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sb->tmem_base;
    const uint32_t ACC_COL0 = PAIR * A_COLS, acc_stride = (uint32_t)((nq + 31) & ~31);
    constexpr uint32_t NACC = MAX_ACC;

    // work items: (a, group of PAIR v-tiles, chunk of captions) -- full groups first, the odd last v-tiles after them, so
    // that the static round-robin deal stays balanced; tiles inside an item: (b, q-tile, v-tile of the group).
    // REDUCE: an item is (a, chunk of captions) and walks ALL v-tile groups of the image as sub-steps, so the running row
    // maxima of the chunk stay in this CTA's shared memory (no global atomics, results written once at the end).
    const int VT = p.VT, QT = p.QT, BCH = p.BCH;
    const int VG = VT / PAIR;  // full groups per image
    const int n_full = p.A * VG * BCH;
    const int n_items = REDUCE ? p.A * BCH : n_full + (VT % PAIR ? p.A * BCH : 0);
    const int n_sub = REDUCE ? (VT + PAIR - 1) / PAIR : 1;
    const int b_per = (p.B + BCH - 1) / BCH;
    auto decode = [&](int item, int sub, int &a, int &vt0, int &ntv, int &b0, int &b1) {
        int bc;
        if (REDUCE) {
            a = item / BCH;
            bc = item - a * BCH;
            vt0 = sub * PAIR;
            ntv = min(PAIR, VT - vt0);
        } else if (item < n_full) {
            a = item / (VG * BCH);
            const int rem = item - a * VG * BCH;
            bc = rem / VG;  // v-tile group fastest: neighbouring CTAs write neighbouring segments of the same rows
            const int g = rem - bc * VG;
            vt0 = g * PAIR;
            ntv = PAIR;
        } else {
            const int j = item - n_full;
            a = j / BCH;
            bc = j - a * BCH;
            vt0 = VG * PAIR;
            ntv = VT - vt0;
        }
        b0 = bc * b_per;
        b1 = min(p.B, b0 + b_per);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t s = 0, s_phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x)
            for (int sub = 0; sub < n_sub; ++sub) {
                int a, vt0, ntv, b0, b1;
                decode(item, sub, a, vt0, ntv, b0, b1);
                for (int t = 0; t < ntv; ++t) {  // image tiles: 2*KB chunks of 16 KB, `cps` per slot
                    const uint8_t *src = p.vis_packed + ((size_t)a * VT + vt0 + t) * (size_t)(2 * KB * CHUNK_A);
                    for (int c = 0; c < 2 * KB; c += cps) {
                        const uint32_t bytes = (uint32_t)min(cps, 2 * KB - c) * CHUNK_A;
                        mbar_wait(&sb->ring_empty[s], s_phase ^ 1);
                        mbar_expect_tx(&sb->ring_full[s], bytes);
                        bulk_g2s(s_ring + (size_t)s * slot_bytes, src + (size_t)c * CHUNK_A, bytes, &sb->ring_full[s]);
                        if (++s == (uint32_t)S) { s = 0; s_phase ^= 1; }
                    }
                }
                for (int b = b0; b < b1; ++b)
                    for (int qt = 0; qt < QT; ++qt) {
                        mbar_wait(&sb->ring_empty[s], s_phase ^ 1);
                        if (p.debug & 8) {  // measurement aid: no caption traffic
                            mbar_arrive(&sb->ring_full[s]);
                        } else {
                            mbar_expect_tx(&sb->ring_full[s], txt_bytes);
                            // the packed caption tile has chunks of 128 rows; copy the first nq rows of each chunk
                            const uint8_t *src = p.txt_packed + ((size_t)b * QT + qt) * (size_t)(2 * KB * CHUNK_A);
                            for (int ch = 0; ch < 2 * KB; ++ch)
                                bulk_g2s(s_ring + (size_t)s * slot_bytes + (size_t)ch * chunk_b, src + (size_t)ch * CHUNK_A,
                                         chunk_b, &sb->ring_full[s]);
                        }
                        if (++s == (uint32_t)S) { s = 0; s_phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // One thread issues every MMA, so its instruction stream must stay well under the tensor time of a tile
        // (24 MMAs x 48 clk): descriptors are formed once per tile and advanced by compile-time constants.
        {   // the whole warp runs this converged; one elected lane issues each tcgen05 instruction
            const uint32_t idesc = idesc_bf16(TILE_M, nq);
            const uint64_t cb4 = (uint64_t)(chunk_b >> 4);  // descriptor units (16 B) between caption chunks
            uint32_t s = 0, s_phase = 0, acc = 0, acc_phase = 0;
            long long t_ring = 0, t_acc = 0, t_vis = 0;
            const long long t_begin = clock64();
            for (int item = blockIdx.x; item < n_items; item += gridDim.x)
            for (int sub = 0; sub < n_sub; ++sub) {
                int a, vt0, ntv, b0, b1;
                decode(item, sub, a, vt0, ntv, b0, b1);
                const long long tv0 = clock64();
                // Image tiles -> tensor memory once per work item. tcgen05.cp executes in issue order behind the MMAs
                // of the previous item, which still read the old tiles. With operand A in tensor memory the MMAs read
                // only the caption operand from shared memory -- with both operands there a 128 x 96 x 16 MMA needs
                // 149 B/clk, more than the 128 B/clk an SM's shared memory delivers.
                for (int t = 0; t < ntv; ++t)
                    for (int c = 0; c < 2 * KB; c += cps) {
                        mbar_wait(&sb->ring_full[s], s_phase);
                        tc_fence_after();
                        const uint64_t desc0 = smem_desc_sw128(smem_u32(s_ring + (size_t)s * slot_bytes));
                        const int n = min(cps, 2 * KB - c);
                        for (int j = 0; j < n; ++j)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_cp_128x256b_elect(tmem_base + (uint32_t)t * A_COLS + (uint32_t)((c + j) * 32 + k * 8),
                                               desc0 + (uint64_t)(j * (CHUNK_A >> 4) + k * 2));
                        tc_commit_elect(&sb->ring_empty[s]);  // the slot may be refilled as soon as the copies retire
                        if (++s == (uint32_t)S) { s = 0; s_phase ^= 1; }
                    }
                const int ntile = (b1 - b0) * QT;
                t_vis += clock64() - tv0;
                for (int tile = 0; tile < ntile; ++tile) {
                    const long long tr0 = clock64();
                    mbar_wait(&sb->ring_full[s], s_phase);
                    t_ring += clock64() - tr0;
                    const uint64_t b_desc0 = smem_desc_sw128(smem_u32(s_ring + (size_t)s * slot_bytes));
                    for (int t = 0; t < ntv; ++t) {
                        const long long ta0 = clock64();
                        mbar_wait(&sb->acc_empty[acc], acc_phase ^ 1);
                        t_acc += clock64() - ta0;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + ACC_COL0 + acc * acc_stride;
                        const uint32_t a_tmem = tmem_base + (uint32_t)t * A_COLS;
                        // hi*hi + lo*hi + hi*lo  (chunk index = part * KB + kb; parts: 0 = hi, 1 = lo)
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const int pa = term == 1 ? 1 : 0, pb = term == 2 ? 1 : 0;
                            if ((term > 0 && p.split == 1) || (p.debug & 4)) break;
#pragma unroll
                            for (int kb = 0; kb < KB; ++kb) {
                                const int ca = pa * KB + kb, cbk = pb * KB + kb;
                                const uint64_t bd = b_desc0 + (uint64_t)cbk * cb4;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K (16 bf16 = 32 B) inside the 128-byte swizzle atom
                                    const uint32_t at = a_tmem + (uint32_t)(ca * 32 + k * 8);
                                    if (term == 0 && kb == 0 && k == 0) tc_mma_ts_elect<0>(d_tmem, at, bd + (uint64_t)(k * 2), idesc);
                                    else tc_mma_ts_elect<1>(d_tmem, at, bd + (uint64_t)(k * 2), idesc);
                                }
                            }
                        }
                        tc_commit_elect(&sb->acc_full[acc]);  // accumulator ready for the epilogue
                        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                    }
                    tc_commit_elect(&sb->ring_empty[s]);  // slot may be refilled once these MMAs have read it
                    if (++s == (uint32_t)S) { s = 0; s_phase ^= 1; }
                }
            }
            if (p.prof && lane == 0) {
                long long *o = p.prof + (size_t)blockIdx.x * 8;
                o[0] = clock64() - t_begin; o[1] = t_ring; o[2] = t_acc; o[3] = t_vis;
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> masks -> (shared ->) global =====================
        // Per (caption tile, image tile) the serial chain of one team -- wait, TMEM -> registers, barrier, registers ->
        // shared, proxy fence, barrier, issue -- is ~1.7k clk, longer than the tensor time of the tile (1.2k clk); two
        // teams on alternate accumulators (each with its own staging tile) take it off the critical path.
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
        const int team = (warp - 2) / kTeamWarps;        // accumulator / staging tile this warp serves
        const int tw = (warp - 2) % kTeamWarps;
        const int half = tw >> 2;                        // which of the 2 warps of the quadrant
        constexpr int MAXCH = 8 / (kEpiWarps / 4);       // 16-query chunks of one warp per tile
        const int T = p.teams;                           // 1: team 0 serves both accumulators (one staging tile fits)
        uint32_t gcount = 0;
        for (int item = blockIdx.x; team < T && item < n_items; item += gridDim.x)
        for (int sub = 0; sub < n_sub; ++sub) {
            int a, vt0, ntv, b0, b1;
            decode(item, sub, a, vt0, ntv, b0, b1);
            bool v_ok_t[PAIR], v_keep_t[PAIR];
#pragma unroll
            for (int t = 0; t < PAIR; ++t) {
                const int vv = (vt0 + t) * VS + quad * 32 + lane;
                v_ok_t[t] = t < ntv && vv < p.V;
                v_keep_t[t] = v_ok_t[t] && p.vis_mask[(size_t)a * p.V + vv] != 0;
            }
            const uint32_t V = (uint32_t)p.ldv;  // row stride of the output (>= V)
            const float neg = p.neg;
            const uint4 *mbp = reinterpret_cast<const uint4 *>(p.txt_maskbits) + (size_t)b0 * QT;
            uint4 mb = mbp[0];  // caption mask bits of the next tile are fetched while the current one is stored
            const size_t tile_rows = (size_t)TILE_M * V, cap_stride = (size_t)p.A * p.Q * V;
            float *ob = p.out + ((size_t)b0 * p.A + a) * p.Q * V + vt0 * VS + quad * 32 + lane;  // (b0, a, q = 0, v)
            for (int b = b0; b < b1; ++b, ob += cap_stride) {
                float *orow0 = ob;
                for (int qt = 0; qt < QT; ++qt, orow0 += tile_rows) {
                    const uint4 mb_cur = mb;
                    ++mbp;
                    if (b + 1 < b1 || qt + 1 < QT) mb = *mbp;
                    const int q_lim = min(TILE_M, p.Q - qt * TILE_M);
#pragma unroll
                    for (int t = 0; t < PAIR; ++t) {
                        if (t >= ntv) break;
                        const uint32_t acc = gcount & 1u, acc_phase = (gcount >> 1) & 1u;
                        ++gcount;
                        if (T == 2 && acc != (uint32_t)team) continue;
                        const bool v_ok = v_ok_t[t], v_keep = v_keep_t[t];
                        float *orow = orow0 + t * VS;
                        mbar_wait(&sb->acc_full[acc], acc_phase);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ACC_COL0 + acc * acc_stride + ((uint32_t)(quad * 32) << 16);
                        if constexpr (!BULK) {
                            // rows that are not 16-byte aligned (odd V without padding): chunk by chunk, one coalesced
                            // 128 B streaming store per (query, 32 factors)
                            for (int c0 = half * 16; c0 < q_lim; c0 += 4 * kEpiWarps) {
                                uint32_t r[16];
                                tc_ld16(taddr + (uint32_t)c0, r);
                                const uint32_t w32 = c0 < 32 ? mb_cur.x : (c0 < 64 ? mb_cur.y : (c0 < 96 ? mb_cur.z : mb_cur.w));
                                const uint32_t w = v_keep ? (w32 >> (c0 & 31)) : 0u;  // bit j = keep the score of query c0 + j
                                float *o = orow + (size_t)c0 * V;
                                if (v_ok && !(p.debug & 1)) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (c0 + j < q_lim) __stcs(o + (uint32_t)j * V, ((w >> j) & 1u) ? __uint_as_float(r[j]) : neg);
                                }
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sb->acc_empty[acc]);
                        } else {
                        // all chunks of this warp TMEM -> registers, then the accumulator goes back to the MMA warp
                        uint32_t r[MAXCH][16];
#pragma unroll
                        for (int k = 0; k < MAXCH; ++k) {
                            const int c0 = half * 16 + k * 4 * kEpiWarps;
                            if (c0 < q_lim && !(p.debug & 2)) tc_ld16_nowait(taddr + (uint32_t)c0, r[k]);
                        }
#pragma unroll
                        for (int k = 0; k < MAXCH; ++k) tc_wait_ld(r[k]);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sb->acc_empty[acc]);
                        if (!v_keep) {  // a masked factor: -INF for every query (per lane, rare)
#pragma unroll
                            for (int k = 0; k < MAXCH; ++k)
#pragma unroll
                                for (int j = 0; j < 16; ++j) r[k][j] = __float_as_uint(neg);
                        }
                        if constexpr (REDUCE) {
                            // (optional) max over the QUERIES, the easy direction of this layout: a thread holds one
                            // factor's scores for its 16-query chunks, so the maximum over q is thread-local; the two
                            // warps of a quadrant (alternate chunks) merge through one shared-memory slot.  First
                            // maximum = smallest q (torch.max); masked queries / factors never win (all masked: neg, 0).
                            float mq = neg;
                            int aq = 0;
                            unsigned long long *s_pq = reinterpret_cast<unsigned long long *>(
                                reinterpret_cast<uint8_t *>(s_run) + p.run_bytes - 2048) + team * TILE_M + quad * 32 + lane;
                            if constexpr (MAXQ) {
#pragma unroll
                                for (int k = 0; k < MAXCH; ++k) {
                                    const int c0 = half * 16 + k * 4 * kEpiWarps;
                                    const uint32_t w32 = k == 0 ? mb_cur.x : (k == 1 ? mb_cur.y : (k == 2 ? mb_cur.z : mb_cur.w));
                                    const uint32_t bits = w32 >> (half * 16);
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (c0 + j < q_lim && ((bits >> j) & 1u)) {
                                            const float x = __uint_as_float(r[k][j]);
                                            if (x > mq) { mq = x; aq = c0 + j; }
                                        }
                                }
                                if (half == 1) *s_pq = ((unsigned long long)__float_as_uint(mq) << 32) | (unsigned)aq;
                            }
                            // transpose through the staged tile, then one warp per query row: each lane takes 4
                            // consecutive factors (one LDS.128), REDUX gives the row maximum, the lowest lane that holds
                            // it gives the first arg-max (torch.max's tie rule); v-tiles merge by atomicMax on
                            // (ordered value bits << 32 | ~v), so equal values keep the smaller factor index
                            float *tile_out = s_out + (size_t)team * out_tile_floats;
                            named_bar(1 + team, 32 * kTeamWarps);  // the previous tile's rows have been reduced
                            if (MAXQ && half == 0) {
                                const unsigned long long o = *s_pq;
                                const float m1 = __uint_as_float((uint32_t)(o >> 32));
                                const int a1 = (int)(uint32_t)o;
                                if (m1 > mq || (m1 == mq && m1 > neg && a1 < aq)) { mq = m1; aq = a1; }
                                if (v_ok) {
                                    const size_t oo = ((size_t)b * p.A + a) * p.V + (size_t)(vt0 + t) * TILE_M + quad * 32 + lane;
                                    p.maxq[oo] = mq;
                                    if (p.argq) p.argq[oo] = aq;
                                }
                            }
#pragma unroll
                            for (int k = 0; k < MAXCH; ++k) {
                                const int c0 = half * 16 + k * 4 * kEpiWarps;
                                float *slot = tile_out + (size_t)c0 * TILE_M + quad * 32 + lane;
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c0 + j < q_lim) slot[j * TILE_M] = __uint_as_float(r[k][j]);
                            }
                            named_bar(1 + team, 32 * kTeamWarps);
                            unsigned long long *dst = s_run + ((size_t)team * b_per + (b - b0)) * p.Q + (size_t)qt * TILE_M;
                            const uint32_t vbase = (uint32_t)((vt0 + t) * TILE_M + lane * 4);
                            for (int q = tw; q < q_lim; q += kTeamWarps) {
                                const uint32_t w32 = q < 32 ? mb_cur.x : (q < 64 ? mb_cur.y : (q < 96 ? mb_cur.z : mb_cur.w));
                                if (!((w32 >> (q & 31)) & 1u)) continue;  // masked query: stays at the initial key
                                const float4 x = *reinterpret_cast<const float4 *>(tile_out + (size_t)q * TILE_M + lane * 4);
                                float m = x.x;
                                uint32_t vi = 0;
                                if (x.y > m) { m = x.y; vi = 1; }
                                if (x.z > m) { m = x.z; vi = 2; }
                                if (x.w > m) { m = x.w; vi = 3; }
                                const uint32_t bits = __float_as_uint(m);
                                const uint32_t key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);  // order-preserving
                                const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
                                const int src = __ffs(__ballot_sync(0xffffffffu, key == kmax)) - 1;
                                const uint32_t vwin = __shfl_sync(0xffffffffu, vbase + vi, src);
                                if (lane == 0) {  // row q of this team's copy belongs to this warp alone: plain update
                                    const unsigned long long k = ((unsigned long long)kmax << 32) | (unsigned long long)(0xffffffffu - vwin);
                                    if (k > dst[q]) dst[q] = k;
                                }
                            }
                        } else {
                            // The tile is staged in shared memory ([query][128 factors] fp32) and written with one
                            // 512 B bulk copy (TMA) per query row: a row segment reaches L2 as one contiguous write and
                            // the warps spend 1 instruction per 4 B on shared memory only. NB tiles alternate, so the
                            // engine drains one while the next is staged.
                            float *tile_out = s_out + (size_t)team * out_tile_floats;
                            float *tile_row0 = orow - (quad * 32 + lane);  // global address of (query 0, first factor of the tile)
                            // word offset modulo 4 of query row r's destination: (k0 + r * vm) & 3
                            const uint32_t k0 = shifted ? (uint32_t)(reinterpret_cast<uintptr_t>(tile_row0) >> 2) & 3u : 0u;
                            const uint32_t vm = shifted ? (V & 3u) : 0u;
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            named_bar(1 + team, 32 * kTeamWarps);  // every issuer's copies out of this tile have been read
#pragma unroll
                            for (int k = 0; k < MAXCH; ++k) {
                                const int c0 = half * 16 + k * 4 * kEpiWarps;
                                float *slot = tile_out + (size_t)c0 * TS + quad * 32 + lane;
                                const uint32_t kc = k0 + (uint32_t)c0 * vm;
                                // the shift of row c0 + j, (kc + j vm) & 3, has period 4 in j: four base pointers, and the 16
                                // stores keep their immediate offsets
                                float *sl[4];
#pragma unroll
                                for (int m = 0; m < 4; ++m) sl[m] = slot + ((kc + (uint32_t)m * vm) & 3u);
                                if (c0 + 16 <= q_lim) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) sl[j & 3][j * TS] = __uint_as_float(r[k][j]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (c0 + j < q_lim) sl[j & 3][j * TS] = __uint_as_float(r[k][j]);
                                }
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            named_bar(1 + team, 32 * kTeamWarps);
                            // a masked query is a row of -INF: its copy reads the constant row instead of the tile,
                            // so the caption mask costs the warps nothing
                            const int row = lane * kTeamWarps + tw;  // <= 16 issuing lanes per warp
                            if (row < q_lim && !(p.debug & 1)) {
                                const uint32_t w32 = row < 32 ? mb_cur.x : (row < 64 ? mb_cur.y : (row < 96 ? mb_cur.z : mb_cur.w));
                                const bool keep = (w32 >> (row & 31)) & 1u;
                                if (!shifted) {
                                    const int len = min(TILE_M, p.ldv - (vt0 + t) * TILE_M);
                                    bulk_s2g(tile_row0 + (size_t)row * V, keep ? tile_out + (size_t)row * TS : s_neg, len * 4);
                                } else {
                                    // factors [x0, x1) of the tile belong to this tile in row `row` (see VSTEP above)
                                    const int tg = vt0 + t;
                                    int x0 = 0, x1 = min(TILE_M, p.ldv - tg * VS);
                                    if (VS != TILE_M) {
                                        constexpr uint32_t U = TILE_M - VS;  // floats per owned unit: a 32-byte sector (8) or a 128-byte line (32)
                                        const uint32_t k8 = (uint32_t)(reinterpret_cast<uintptr_t>(tile_row0) >> 2) & (U - 1);  // VS % U == 0
                                        const int er = (int)((U - ((k8 + (uint32_t)row * (V & (U - 1))) & (U - 1))) & (U - 1));
                                        if (tg > 0) x0 = er;
                                        if (tg < VT - 1) x1 = VS + er;
                                    }
                                    const int len = x1 - x0;
                                    const int kr = (int)((k0 + (uint32_t)row * vm) & 3u);
                                    const int h = (4 - ((kr + x0) & 3)) & 3;                                       // floats up to alignment
                                    float *dst = tile_row0 + (size_t)row * V + x0;
                                    const float *src = tile_out + (size_t)row * TS + kr + x0;                      // element e at src[e]
                                    const int nb = len > h ? ((len - h) & ~3) : 0;
                                    if (nb > 0) bulk_s2g(dst + h, keep ? src + h : s_neg, nb * 4);
                                    for (int e = 0; e < min(h, len); ++e) __stcs(dst + e, keep ? src[e] : neg);
                                    for (int e = h + nb; e < len; ++e) __stcs(dst + e, keep ? src[e] : neg);
                                }
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        }
                    }
                }
            }
            if (REDUCE && sub == n_sub - 1) {
                // every v-tile of the image has been seen for this chunk of captions: merge the teams' copies, unpack and
                // write max / arg-max once (masked queries were never touched: fill value, index 0)
                named_bar(3, 32 * kTeamWarps * T);
                const int e_tid = (warp - 2) * 32 + lane, e_n = 32 * kTeamWarps * T, cnt = (b1 - b0) * p.Q;
                for (int e = e_tid; e < cnt; e += e_n) {
                    const int bb = e / p.Q, q = e - bb * p.Q;
                    unsigned long long k = s_run[(size_t)bb * p.Q + q];
                    s_run[(size_t)bb * p.Q + q] = 0ull;
                    if (T == 2) {
                        const unsigned long long k1 = s_run[((size_t)b_per + bb) * p.Q + q];
                        s_run[((size_t)b_per + bb) * p.Q + q] = 0ull;
                        k = k1 > k ? k1 : k;
                    }
                    float m = p.neg;
                    int arg = 0;
                    if (k != 0ull) {
                        const uint32_t key = (uint32_t)(k >> 32);
                        m = __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
                        arg = (int)(0xffffffffu - (uint32_t)k);
                    }
                    const size_t o = ((size_t)(b0 + bb) * p.A + a) * p.Q + q;
                    p.maxv[o] = m;
                    if (p.argv) p.argv[o] = arg;
                }
                named_bar(3, 32 * kTeamWarps * T);
            }
        }
        if (BULK && !REDUCE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

int g_align_sm = 0, g_align_smem = 0;
cudaError_t align_device_info() {
    if (g_align_sm) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_align_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    return cudaDeviceGetAttribute(&g_align_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}

}  // namespace

AlignPlan align_plan(int A, int V, int B, int Q, int D, int vstep) {
    AlignPlan pl{};
    pl.KB = (D + 63) / 64;
    // tiles of 128 factors whose first factors are vstep apart (128: disjoint; 120: overlapping, align_gemm_kernel VSTEP)
    pl.VT = V <= TILE_M ? 1 : (V - (TILE_M - vstep) + vstep - 1) / vstep;
    pl.QT = (Q + TILE_M - 1) / TILE_M;
    const int qmax = Q < TILE_M ? Q : TILE_M;
    pl.nq = (qmax + 15) & ~15;
    pl.tile_bytes = (size_t)2 * pl.KB * CHUNK_A;
    pl.vis_packed_bytes = (size_t)A * pl.VT * pl.tile_bytes;
    pl.txt_packed_bytes = (size_t)B * pl.QT * pl.tile_bytes;
    pl.maskbits_bytes = (size_t)B * pl.QT * 4 * sizeof(uint32_t);
    return pl;
}

size_t align_workspace_bytes(int A, int V, int B, int Q, int D) {
    const AlignPlan pl = align_plan(A, V, B, Q, D), pd = align_plan(A, V, B, Q, D, TILE_M - 32);  // room for any tiling
    return (pl.vis_packed_bytes > pd.vis_packed_bytes ? pl.vis_packed_bytes : pd.vis_packed_bytes) + pl.txt_packed_bytes +
           ((pl.maskbits_bytes + 255) & ~(size_t)255) + 1024;
}

// pack both operands (bf16 hi / lo, swizzled tile images) + the caption mask bits into the workspace:
//   [vis tiles][caption tiles][mask bits]   (shared by the forward and the backward kernels)
cudaError_t align_pack_operands(const float *vis, const float *txt, const uint8_t *txt_mask, int A, int V, int B, int Q, int D,
                                void *workspace, cudaStream_t st, int vstep) {
    cudaError_t e = align_device_info();
    if (e != cudaSuccess) return e;
    const AlignPlan pl = align_plan(A, V, B, Q, D, vstep);
    uint8_t *ws = reinterpret_cast<uint8_t *>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    uint8_t *vis_packed = ws;
    uint8_t *txt_packed = vis_packed + pl.vis_packed_bytes;
    uint32_t *maskbits = reinterpret_cast<uint32_t *>(txt_packed + pl.txt_packed_bytes);
    const size_t nv = (size_t)A * pl.VT * TILE_M * pl.KB * 8, nt = (size_t)B * pl.QT * TILE_M * pl.KB * 8;
    int gv = (int)((nv + 255) / 256), gt = (int)((nt + 255) / 256);
    const int cap = g_align_sm * 16;
    if (gv > cap) gv = cap;
    if (gt > cap) gt = cap;
    align_pack_kernel<<<gv, 256, 0, st>>>(vis, nullptr, A, V, D, pl.KB, TILE_M, pl.tile_bytes, pl.VT, vis_packed, nullptr, vstep);
    align_pack_kernel<<<gt, 256, 0, st>>>(txt, txt_mask, B, Q, D, pl.KB, TILE_M, pl.tile_bytes, pl.QT, txt_packed, maskbits, TILE_M);
    return cudaGetLastError();
}

// mode 0 / 1: materialise the logits in `out`; mode 2: reduce over the factors (maxv / argv)
static cudaError_t launch_align_mode(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask,
                                     int A, int V, int B, int Q, int D, float neg, int split, float *out, int ldv,
                                     float *maxv, int *argv, void *workspace, cudaStream_t st, float *maxq = nullptr,
                                     int *argq = nullptr) {
    cudaError_t e = align_device_info();
    if (e != cudaSuccess) return e;
    const bool reduce = maxv != nullptr;
    // bulk (TMA) stores need 16-byte aligned row segments; otherwise the warps store directly
    // (rows that are not 16-byte aligned -- odd V without padding, the reference's own layout -- use the shifted staging,
    // with overlapping tiles 120 factors apart so that whole 32-byte sectors are written; VLGAE_ALIGN_VSTEP=128 for A/B runs)
    const bool aligned_rows = (ldv % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    bool bulk;
    { const char *bk = getenv("VLGAE_ALIGN_BULK"); bulk = reduce || (bk ? atoi(bk) != 0 : true); }
    const bool shift = bulk && !(reduce || aligned_rows);
    int vstep = TILE_M;
    if (shift) { const char *vs = getenv("VLGAE_ALIGN_VSTEP"); vstep = (vs && atoi(vs) == TILE_M) ? TILE_M : TILE_M - 8; }
    const AlignPlan pl = align_plan(A, V, B, Q, D, vstep);
    uint8_t *ws = reinterpret_cast<uint8_t *>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    uint8_t *vis_packed = ws;
    uint8_t *txt_packed = vis_packed + pl.vis_packed_bytes;
    uint32_t *maskbits = reinterpret_cast<uint32_t *>(txt_packed + pl.txt_packed_bytes);

    e = align_pack_operands(vis, txt, txt_mask, A, V, B, Q, D, workspace, st, vstep);
    if (e != cudaSuccess) return e;

    AlignArgs a{};
    a.vis_packed = vis_packed; a.txt_packed = txt_packed; a.txt_maskbits = maskbits; a.vis_mask = vis_mask;
    a.out = out; a.ldv = ldv; a.A = A; a.V = V; a.B = B; a.Q = Q; a.KB = pl.KB; a.VT = pl.VT; a.QT = pl.QT; a.nq = pl.nq;
    a.neg = neg; a.split = split == 1 ? 1 : 3;
    { const char *dbg = getenv("VLGAE_ALIGN_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }  // measurement aids, see the kernel
    a.prof = dmv_profile_buffer();
    a.maxv = maxv; a.argv = argv; a.maxq = maxq; a.argq = argq;
    if (maxq && pl.QT != 1) return cudaErrorInvalidValue;  // the max over the queries lives inside one query tile
    a.bulk = bulk;
    a.tile_stride = (reduce || aligned_rows) ? TILE_M : TILE_M + 4;
    // shared memory: ring slots (a caption tile, or chunks of an image tile) + staged output tiles
    size_t slot_bytes = (size_t)2 * pl.KB * pl.nq * 128;
    if (slot_bytes < (size_t)CHUNK_A) slot_bytes = CHUNK_A;
    const int out_rows = Q < pl.nq ? Q : pl.nq;
    // REDUCE: captions per work item bounded by the running-maxima arrays (2 teams x b_per x Q x 8 B <= 24 KB)
    int bch = 1, b_per = B;
    size_t run_bytes = 0;
    if (reduce) {
        int bmax = (int)(24576 / ((size_t)16 * Q));
        if (bmax < 1) bmax = 1;
        bch = (B + bmax - 1) / bmax;
        while ((long long)A * bch < 6LL * g_align_sm && bch < B) ++bch;
        b_per = (B + bch - 1) / bch;
        run_bytes = (((size_t)2 * b_per * Q * 8 + 15) & ~(size_t)15) + 2048;  // + exchange slots of the max over the queries
    }
    const size_t out_tile_bytes = (size_t)out_rows * a.tile_stride * 4, fixed = sizeof(AlignSmem) + 64 + 512 + run_bytes;
    int out_bufs = a.bulk ? 2 : 0;
    if (out_bufs == 2 && ((size_t)g_align_smem - fixed - 2 * out_tile_bytes) / slot_bytes < 2) out_bufs = 1;
    int stages = (int)(((size_t)g_align_smem - fixed - out_bufs * out_tile_bytes) / slot_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorInvalidValue;
    a.stages = stages; a.out_bufs = out_bufs; a.slot_bytes = (uint32_t)slot_bytes;
    a.teams = a.bulk ? out_bufs : kTeams; a.out_rows = out_rows; a.run_bytes = (uint32_t)run_bytes;
    const size_t smem_bytes = (size_t)stages * slot_bytes + out_bufs * out_tile_bytes + 512 + run_bytes + sizeof(AlignSmem) + 64;
    // Work items are dealt round-robin to the persistent CTAs: split the captions of one (image, v-tile pair) into
    // chunks until every CTA gets >= 16 items, so the uneven last round costs a few per cent at most.
    const long long groups = reduce ? (long long)A : (long long)A * ((pl.VT + 1) / 2);
    if (!reduce) {
        while (groups * bch < 16LL * g_align_sm && bch < B) bch <<= 1;
        if (bch > B) bch = B;
    }
    a.BCH = bch;
    int grid = g_align_sm;
    if (grid > groups * bch) grid = (int)(groups * bch);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (err != cudaSuccess) return err;
        kern<<<grid, kThreads, smem_bytes, st>>>(a);
        return cudaGetLastError();
    };
    if (reduce && a.maxq) return pl.KB == 1 ? launch(align_gemm_kernel<1, 3>) : launch(align_gemm_kernel<2, 3>);
    if (reduce) return pl.KB == 1 ? launch(align_gemm_kernel<1, 2>) : launch(align_gemm_kernel<2, 2>);
    if (shift && vstep == TILE_M) return pl.KB == 1 ? launch(align_gemm_kernel<1, 1, TILE_M>) : launch(align_gemm_kernel<2, 1, TILE_M>);
    if (shift) return pl.KB == 1 ? launch(align_gemm_kernel<1, 1, TILE_M - 8>) : launch(align_gemm_kernel<2, 1, TILE_M - 8>);
    if (pl.KB == 1) return a.bulk ? launch(align_gemm_kernel<1, 1>) : launch(align_gemm_kernel<1, 0>);
    return a.bulk ? launch(align_gemm_kernel<2, 1>) : launch(align_gemm_kernel<2, 0>);
}

cudaError_t launch_align(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                         int V, int B, int Q, int D, float neg, int split, float *out, int ldv, void *workspace,
                         cudaStream_t st) {
    return launch_align_mode(vis, vis_mask, txt, txt_mask, A, V, B, Q, D, neg, split, out, ldv, nullptr, nullptr, workspace, st);
}

size_t align_reduce_bytes(int, int, int) { return 0; }  // the running maxima live in shared memory

cudaError_t launch_align_reduce(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                                int V, int B, int Q, int D, float neg, int split, float *maxv, int *argv, void *workspace,
                                cudaStream_t st) {
    return launch_align_mode(vis, vis_mask, txt, txt_mask, A, V, B, Q, D, neg, split, nullptr, 0, maxv, argv, workspace, st);
}

cudaError_t launch_align_maxima(const float *vis, const uint8_t *vis_mask, const float *txt, const uint8_t *txt_mask, int A,
                                int V, int B, int Q, int D, float neg, int split, float *maxv, int *argv, float *maxq,
                                int *argq, void *workspace, cudaStream_t st) {
    return launch_align_mode(vis, vis_mask, txt, txt_mask, A, V, B, Q, D, neg, split, nullptr, 0, maxv, argv, workspace, st, maxq,
                             argq);
}

}  // namespace vlgae
