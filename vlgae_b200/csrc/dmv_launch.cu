// dmv_launch.cu -- host-side launch logic of the DMV kernels + the small element-wise kernels of the path.
//
// The chart kernels themselves live in dmv_frontier.cu (latency regime, zero-copy hand-off, charts beyond shared
// memory) and dmv_gather.cu (gather schedule).  This file chooses the schedule, CTA size and length buckets per launch
// and holds DMV1o.merge (reference /root/reference/src/model/torch_struct/distributions.py:253-265), the row scaling
// used by the autograd nodes and the two roofline microbenchmarks.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dmv_kernels.cuh"

namespace vlgae {

namespace {

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
static int g_sm_count = 0, g_smem_optin = 0, g_info_dev = -1;
static cudaError_t device_info() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev == g_info_dev) return cudaSuccess;
    e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e == cudaSuccess) g_info_dev = dev;
    return e;
}

int dmv_sm_count() { return device_info() == cudaSuccess ? g_sm_count : 148; }

size_t dmv_ws_slice_bytes(int N, int passes) { return dmv_frontier_chart_bytes(N, passes); }

bool dmv_fits_smem(int N, int passes) {
    if (device_info() != cudaSuccess) return false;
    return dmv_frontier_fits(N, passes, g_smem_optin);
}

int dmv_grid_for_workspace(int B) {
    if (device_info() != cudaSuccess) return 0;
    const int cap = g_sm_count * 4;
    return B * 2 < cap ? B * 2 : cap;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
static int g_schedule = 0;  // 0 = automatic, 1 = frontier, 2 = gather
void dmv_set_schedule(int which) { g_schedule = which; }
static int g_lin_max_len = -1;  // < 0: the default (VLGAE_FRONTIER_LINEAR, else 24)
void dmv_set_linear_max_len(int words) { g_lin_max_len = words; }
static long long *g_prof = nullptr;
void dmv_set_profile_buffer(long long *buf) { g_prof = buf; }
long long *dmv_profile_buffer() { return g_prof; }

// frontier launch configuration for a length bucket sized for `cap` positions
static cudaError_t launch_frontier_cap(const DmvArgs &a, int passes, int cap, cudaStream_t st) {
    static const int env_ft = env_int("VLGAE_FRONTIER_THREADS", 0);
    const bool fits = dmv_frontier_fits(cap, passes, g_smem_optin);
    const bool resident = (long long)a.B * a.npass <= 2LL * g_sm_count && !a.only;
    if (cap > 256 || !(fits || a.workspace)) return cudaErrorInvalidValue;
    DmvArgs f = a;
    // Latency regime (every work item resident at once): many threads, running state in registers.  Throughput
    // regime: small CTAs with the state in shared memory (idle warps skip a phase entirely) -- measured on B200:
    // COCO-like bulk 806 us vs 1126, 512 x 16 words 31 us vs 58; 40-word charts prefer 256 threads / registers.
    // Charts beyond shared memory (N > ~72): same kernel, chart arrays in the CTA's workspace slice (L2-resident).
    int ft;
    bool reg_state;
    if (!fits) { ft = env_int("VLGAE_FRONTIER_BIG_THREADS", 1024); reg_state = false; }
    else if (resident) { ft = cap <= 24 ? 256 : 512; reg_state = true; }
    // warp per sentence for short charts; with 5 cells per lane (15-17 positions) a sentence takes longer than in a
    // 128-thread CTA, so that size only switches inside bulk (length-bucketed) launches (682 vs 714 us)
    else if (cap <= 14 || (cap <= 17 && a.nb_hi < a.N))  // nb_hi < N: a length bucket of a bulk launch
        { ft = env_int("VLGAE_FRONTIER_WARP", 1) ? 32 : (cap <= 12 ? 64 : 128); reg_state = false; }
    // 15..23 positions: two cells per thread in registers once the linear-domain sweeps apply to every sentence of the launch
    // (512 x 16 words: 15.8 -> 18.0 M sentences/s; in the log domain the shared-memory state was faster: 31 vs 58 us)
    else if (cap <= 23 && a.lin_max_len >= cap - 1) { ft = 128; reg_state = true; }
    else if (cap <= 33) { ft = 128; reg_state = false; }
    else if (cap <= 45) { ft = 256; reg_state = true; }   // <= 1024 cells: 4 per thread in registers
    else { ft = 512; reg_state = env_int("VLGAE_FRONTIER_REG5", 1) != 0; }  // one CTA per SM: more threads (n = 64: 907 vs 1039 us);
                                                          // <= 2560 cells: five per thread in registers
    if (env_ft > 0) ft = env_ft;
    if (fits) { f.workspace = nullptr; f.ws_stride = 0; }
    else f.ws_stride = dmv_ws_slice_bytes(a.N, 3);
    return launch_dmv_frontier(f, passes, cap, ft, reg_state, g_sm_count, dmv_grid_for_workspace(a.B), st);
}

// Independent launches of one call (the two semirings, the length ranges / buckets: disjoint sentences or disjoint
// outputs) go to side streams forked from the caller's stream and joined before the call returns to it, so the tail of
// one launch (its last, partly empty wave) overlaps the head of the next instead of idling the SMs.  Event fork / join
// only: legal under stream capture.  VLGAE_DMV_LANES=0 keeps everything on the caller's stream.
struct Lanes {
    static constexpr int NSIDE = 6;
    cudaStream_t main = nullptr, side[NSIDE] = {};
    cudaEvent_t fork = nullptr, join[NSIDE] = {};
    bool used[NSIDE] = {};
    bool on = false;
    int dev = -1;
    cudaError_t begin(cudaStream_t st, bool enable) {
        main = st; on = false;
        for (bool &u : used) u = false;
        if (!enable) return cudaSuccess;
        int d = 0;
        cudaError_t e = cudaGetDevice(&d);
        if (e != cudaSuccess) return e;
        if (d != dev) {  // first use on this device (streams of a previous device are left to the driver)
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            static const int env_prio = env_int("VLGAE_DMV_LANE_PRIORITY", 1);
            for (int k = 0; k < NSIDE; ++k) {
                // lanes 1..3 carry the log-semiring launches, the long ones: their CTAs are placed first
                const int prio = (k < 3 && env_prio) ? greatest : least;
                if ((e = cudaStreamCreateWithPriority(&side[k], cudaStreamNonBlocking, prio)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&join[k], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
            if ((e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)) != cudaSuccess) return e;
            dev = d;
        }
        if ((e = cudaEventRecord(fork, st)) != cudaSuccess) return e;
        on = true;
        return cudaSuccess;
    }
    // lane 0 is the caller's stream
    cudaStream_t get(int lane) {
        if (!on || lane <= 0) return main;
        const int k = (lane - 1) % NSIDE;
        if (!used[k]) { cudaStreamWaitEvent(side[k], fork, 0); used[k] = true; }
        return side[k];
    }
    cudaError_t end() {
        if (!on) return cudaSuccess;
        for (int k = 0; k < NSIDE; ++k) {
            if (!used[k]) continue;
            cudaError_t e = cudaEventRecord(join[k], side[k]);
            if (e != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(main, join[k], 0)) != cudaSuccess) return e;
        }
        on = false;
        return cudaSuccess;
    }
};
static thread_local Lanes g_lanes;

// Gather schedule (dmv_gather.cu: linear-domain log semiring, value-only Viterbi) for the sentences with lo <= len + 1 <= hi.
// With `split`, one launch pair per row stride of the layout (charts of <= 25 / 33 / 41 positions: 10 / 6 / 4
// log-semiring CTAs per SM), longest first; the log-semiring launches take the high-priority lanes 1..3, the Viterbi
// launches lane 0 (the caller's stream) and lanes 4, 5.
// The caller joins the lanes and then calls launch_gather_redo: ONE frontier launch restricted to the sentences whose
// linear-domain sweep flagged itself (normally none: that launch then costs a few microseconds of flag reads).
static cudaError_t launch_gather_ranges(const DmvArgs &a, int passes, int lo, int hi, bool split, Lanes &lanes) {
    static const int env_gt = env_int("VLGAE_GATHER_THREADS", 0);
    static const int ghi[] = {41, 33, 25}, glo[] = {34, 26, 0};
    for (int which = 1; which <= 2; ++which) {  // log-semiring launches first: the longer ones
        if (!(passes & which)) continue;
        for (int k = 0; k < 3; ++k) {
            int rlo = glo[k] > lo ? glo[k] : lo, rhi = ghi[k] < hi ? ghi[k] : hi;
            if (!split) { rlo = lo; rhi = hi; }
            if (rlo > rhi) continue;
            DmvArgs f = a;
            f.workspace = nullptr; f.ws_stride = 0; f.only = nullptr;
            f.nb_lo = rlo; f.nb_hi = rhi;
            const int lane = which == 1 ? 1 + k : (k == 0 ? 0 : 3 + k);
            if (a.counter) f.counter = a.counter + (which - 1) * 3 + k;
            cudaError_t e = launch_dmv_gather(f, which, rhi, env_gt, g_sm_count, lanes.get(lane));
            if (e != cudaSuccess) return e;
            if (!split) break;
        }
    }
    return cudaSuccess;
}
static cudaError_t launch_gather_redo(const DmvArgs &a, int passes, int lo, int hi, cudaStream_t st) {
    if (!(passes & 1)) return cudaSuccess;
    DmvArgs r = a;
    r.npass = 1; r.first_pass = 0; r.only = a.redo; r.redo = nullptr; r.workspace = nullptr; r.ws_stride = 0;
    r.nb_lo = lo; r.nb_hi = hi;
    r.lin_max_len = 0; r.lin_long_from = 0;  // these sentences failed the linear-domain self-check once already: log domain straight away
    return launch_frontier_cap(r, 1, hi, st);
}

static bool gather_usable(const DmvArgs &a, int passes, int cap) {
    return a.redo && !a.share && !a.only && dmv_gather_fits(cap, passes, g_smem_optin);
}

cudaError_t launch_dmv(const DmvArgs &a_in, int passes, cudaStream_t st) {
    DmvArgs a = a_in;
    a.prof = g_prof;
    cudaError_t e = device_info();
    if (e != cudaSuccess) return e;
    a.nsm = g_sm_count;
    a.nb_lo = 0; a.nb_hi = a.N;
    static const int env_no_off = env_int("VLGAE_DMV_NO_OFFSETS", 0);
    a.no_offsets = env_no_off;
    // Linear-domain frontier sweeps (dmv_frontier.cu, LIN) for sentences of <= 24 words by default.  The variant is faster
    // (512 x 8 words: 19.0 -> 27.0 M sentences/s, cfg1 +9 %, 100k captions +4 %; 40 words: 53.3 -> 49.4 us) and closer to the
    // exact marginals everywhere (2.5e-7 instead of 1.3e-6 from fp64 on the cfg2 batch) -- and exactly that breaks the plain
    // 1e-5 gate on LONG sentences: the reference's own fp32 sweep is 1.0012e-5 from fp64 on the cfg2 batch (its error grows
    // with |log Z| ~ 4 len: one ulp is 1.5e-5 from 128 on), so the closer result lands 1.0014e-5 from the REFERENCE where the
    // log-domain sweep happens to land at 9.78e-6.  Up to 24 words |log Z| < 128 keeps the reference within ~5e-6 of fp64 and
    // both arithmetics within the tolerance; beyond (25..44 words), parity with the reference comes first.
    // VLGAE_FRONTIER_LINEAR = 0: never, 1: every length, n > 1: up to n words.
    static const int env_lin = env_int("VLGAE_FRONTIER_LINEAR", 24);
    a.lin_max_len = g_lin_max_len >= 0 ? g_lin_max_len : (env_lin == 1 ? 1 << 20 : env_lin);
    // ... and again from 45 words on (the five-cells-per-thread charts of 46..72 positions): there the reference's own fp32
    // result is > 1e-5 from the exact one (1.4e-5 at n = 64), so parity is the three-way rule either way, which the more exact
    // sweep always meets; n = 64 x 512: 711 k -> 804 k sentences/s.  Off together with the short range (0).
    a.lin_long_from = a.lin_max_len > 0 ? 45 : 0;
    static const int env_prof_all = env_int("VLGAE_PROF_ALL", 0);
    a.prof_all = env_prof_all;
    static const int env_retry = env_int("VLGAE_DMV_RETRY_ABOVE", 0);
    a.retry_above = (float)env_retry;
    static const int env_sched = [] {
        const char *v = getenv("VLGAE_DMV_KERNEL");
        return !v ? 0 : (v[0] == 'f' ? 1 : (v[0] == 'g' ? 2 : 0));
    }();
    const int sched = g_schedule ? g_schedule : env_sched;
    // Throughput regime (more work items than one resident wave at the padded length): one launch per length bucket,
    // shared memory sized for the bucket, so short sentences run at 10-20 CTAs per SM instead of the 3 a 40-word
    // chart allows.  Sentences outside a launch's bucket are skipped by its CTAs (no host knowledge of the lengths,
    // no sorting assumption).  Longest bucket first.
    static const int env_bucket = env_int("VLGAE_DMV_BUCKETS", 1);
    // measured on B200 (tools/dmv_sweep.py, 512 full-length sentences): the gather schedule overtakes the frontier
    // schedule between 24 and 28 words (n = 24: 75 vs 52 us, n = 28: 89 vs 103 us, n = 40: 157 vs 220 us)
    static const int env_gather_lo = env_int("VLGAE_GATHER_MIN_POSITIONS", 28);
    // in a bulk launch (length buckets, tens of waves) the gather schedule already wins from 18 positions on
    // (COCO-shaped 16384 sentences: 585 us with the buckets >= 18 on the gather schedule, 661 us with those >= 28)
    // (r2c, lanes + length split: 528 us from 12 positions on, 604 us from 18; re-tuned on 100 k COCO-shaped captions after
    // the one-phase forward sweeps of the frontier kernel: 12 / 14 / 16 / 18 / 20 -> 3156 / 3008 / 3206 / 3392 / 3142 us)
    static const int env_gather_bulk_lo = env_int("VLGAE_GATHER_BULK_MIN_POSITIONS", 14);
    static const int env_lanes = env_int("VLGAE_DMV_LANES", 1);
    static const int env_split = env_int("VLGAE_GATHER_SPLIT", 1);
    static const int env_dynamic = env_int("VLGAE_GATHER_DYNAMIC", 1);
    const long long items = (long long)a.B * a.npass;
    const bool bulk = env_bucket && items > 8192 && !a.share;
    const int gcap = a.N < DMV_GATHER_MAX_POSITIONS ? a.N : DMV_GATHER_MAX_POSITIONS;
    if (!bulk) {
        // one launch.  Latency regime (every work item resident at once): frontier schedule.  Beyond one wave the
        // gather schedule takes the batches padded to env_gather_lo .. 41 positions (and any batch on request: tests,
        // sweeps) -- lengths are not known on the host, so the padded length decides.
        const bool resident = items <= 2LL * g_sm_count;
        const bool want = sched == 2 || (sched == 0 && !resident && a.N >= env_gather_lo);
        // (a batch of several waves is split by length like a bulk launch: short sentences then run at 6-10 CTAs per SM)
        if (want && a.N <= DMV_GATHER_MAX_POSITIONS && gather_usable(a, passes, a.N)) {
            Lanes &lanes = g_lanes;
            if (!env_dynamic) a.counter = nullptr;
            if (a.counter && (e = cudaMemsetAsync(a.counter, 0, DMV_GATHER_COUNTERS * sizeof(int), st)) != cudaSuccess) return e;
            if ((e = lanes.begin(st, env_lanes && passes == 3 && a.counter)) != cudaSuccess) return e;
            // (one launch pair: splitting a few waves by length adds more launch tails than the residency gains)
            e = launch_gather_ranges(a, passes, 0, a.N, env_split > 1, lanes);
            const cudaError_t e2 = lanes.end();
            if (e != cudaSuccess || e2 != cudaSuccess) return e != cudaSuccess ? e : e2;
            return launch_gather_redo(a, passes, 0, a.N, st);
        }
        return launch_frontier_cap(a, passes, a.N, st);
    }
    Lanes &lanes = g_lanes;
    if (!env_dynamic) a.counter = nullptr;
    if (a.counter && (e = cudaMemsetAsync(a.counter, 0, DMV_GATHER_COUNTERS * sizeof(int), st)) != cudaSuccess) return e;
    if ((e = lanes.begin(st, env_lanes != 0)) != cudaSuccess) return e;
    int gather_lo = a.N + 1;  // lowest chart size handled by the gather launches
    if (sched != 1 && gather_usable(a, passes, gcap) && env_gather_bulk_lo <= gcap) {
        gather_lo = env_gather_bulk_lo;
        e = launch_gather_ranges(a, passes, gather_lo, gcap, env_split != 0, lanes);
        if (e != cudaSuccess) { lanes.end(); return e; }
    }
    const int gather_hi = gather_lo <= a.N ? gcap : 0;  // [gather_lo, gather_hi] is done
    static const int caps[] = {8, 12, 16, 20, 24, 28, 33, 41, 49, 65, 97, 129, 256};
    int nb = 0, bounds[16];
    for (int c : caps) if (c < a.N) bounds[nb++] = c;
    bounds[nb++] = a.N;
    for (int k = nb - 1; k >= 0; --k) {
        DmvArgs bkt = a;
        bkt.nb_hi = bounds[k];
        bkt.nb_lo = k > 0 ? bounds[k - 1] + 1 : 0;
        // clip the bucket against the range the gather launches took
        if (gather_hi) {
            if (bkt.nb_lo >= gather_lo && bkt.nb_hi <= gather_hi) continue;
            if (bkt.nb_lo < gather_lo && bkt.nb_hi >= gather_lo) bkt.nb_hi = gather_lo - 1;
            else if (bkt.nb_lo <= gather_hi && bkt.nb_hi > gather_hi) bkt.nb_lo = gather_hi + 1;
        }
        // the frontier buckets share one lane (each is a short launch; the lane runs beside the gather launches)
        e = launch_frontier_cap(bkt, passes, bkt.nb_hi, lanes.get(gather_hi ? 6 : 0));
        if (e != cudaSuccess) { lanes.end(); return e; }
    }
    if ((e = lanes.end()) != cudaSuccess) return e;
    if (gather_hi) return launch_gather_redo(a, passes, gather_lo, gather_hi, st);
    return cudaSuccess;
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
