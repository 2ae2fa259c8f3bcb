// tools/probes/bulk_store.cu -- write the alignment output [B,A,Q,ldv] with 1-D bulk TMA stores from shared memory:
// each persistent CTA owns a 128-float column strip of one image and stores Q rows x 512 B per caption.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/bulk_store tools/probes/bulk_store.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: every row by lanes of warp 0 (32 copies in flight per round); mode 1: rows spread over all warps' lane 0
__global__ void __launch_bounds__(256) bulk_writer(float *out, int A, int B, int Q, int ldv, int nstrip, int nbuf, int mode, int W) {
    extern __shared__ __align__(128) float tile[];  // nbuf x [Q][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nbuf * Q * W; i += blockDim.x) tile[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int n_items = A * nstrip;
    int buf = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int a = item / nstrip, s = item - a * nstrip;
        const int wbytes = min(W, ldv - s * W) * 4;
        for (int b = 0; b < B; ++b) {
            float *base = out + ((size_t)(b * A + a) * Q) * ldv + (size_t)s * W;
            const float *src = tile + (size_t)buf * Q * W;
            if (mode == 0) {
                if (warp == 0) {
                    for (int q = lane; q < Q; q += 32)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)q * ldv),
                                     "r"(smem_u32(src + q * W)), "r"(wbytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
            } else if (mode >= 2) {
                const int qb = mode;  // rows [0, qb) by bulk copies, [qb, Q) by the warps (float4 per lane, 512 B per warp)
                for (int q = threadIdx.x; q < qb; q += blockDim.x)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)q * ldv),
                                 "r"(smem_u32(src + q * W)), "r"(wbytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                for (int q = qb + warp; q < Q; q += 8) {
                    const float4 v = reinterpret_cast<const float4 *>(src + q * W)[lane];
                    if (lane * 16 < wbytes) __stcs(reinterpret_cast<float4 *>(base + (size_t)q * ldv) + lane, v);
                }
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            } else {
                for (int q = threadIdx.x; q < Q; q += blockDim.x)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)q * ldv),
                                 "r"(smem_u32(src + q * W)), "r"(wbytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncthreads();  // stands in for the epilogue refilling the buffer
            if (++buf == nbuf) buf = 0;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int A = 128, B = 128, Q = 82, ldv = 1376;
    const size_t n = (size_t)A * B * Q * ldv;
    float *out;
    cudaMalloc(&out, n * 4);
    cudaMemset(out, 0, n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(bulk_writer, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int qb = 2; qb <= 82; qb += 10) {
        const int W = 128, nbuf = 2, grid = 148;
        float ms = 0;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            bulk_writer<<<grid, 256, (size_t)nbuf * Q * W * 4>>>(out, A, B, Q, ldv, 11, nbuf, qb, W);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("rows by bulk copy %2d / by warps %2d: %.3f ms  %.0f GB/s (%s)\n", qb, Q - qb, ms, n * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    // check every element was written
    float *h = (float *)malloc(1 << 20);
    cudaMemcpy(h, out + n - (1 << 18), 1 << 20, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int i = 0; i < (1 << 18); ++i) bad += h[i] != 1.0f;
    printf("tail check: %zu unwritten of %d\n%s\n", bad, 1 << 18, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
