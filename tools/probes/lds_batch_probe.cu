// lds_batch_probe.cu -- how long does one warp take for a batch of shared-memory loads + FMAs (the gather kernel's
// inner step), as a function of the loads per batch, their width and the number of warps per SM doing the same?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_batch_probe lds_batch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NB>  // terms per batch; per term: 4 x LDS.64 + 2 x LDS.32 + 5 FMA
__global__ void probe(float *out, long long *clk, int iters, int stride) {
    extern __shared__ float2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 6000; t += blockDim.x) sm[t] = make_float2(1.0f + t * 1e-6f, 0.5f);
    __syncthreads();
    const float2 *pL = sm + lane * stride + warp * 7, *pR = pL + 500, *iL = pL + 1800, *iR = pL + 2300;
    const float *tL = reinterpret_cast<const float *>(sm + 4000) + lane * stride + warp * 3, *tR = tL + 777;
    float2 x = make_float2(0.f, 0.f), cl = x, cr = x;
    int kk = clk[1] & 31;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float2 vL[NB], vR[NB], v3[NB], v4[NB];
        float s3[NB], s4[NB];
        kk = (kk * 5 + 3) & 31; const int k = kk;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            vL[u] = pL[k + u]; vR[u] = pR[k + u]; s3[u] = tL[k + u]; s4[u] = tR[k + u];
            v3[u] = iL[k + u]; v4[u] = iR[k + u];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            x.x = fmaf(vL[u].x, vR[u].x, x.x); x.y = fmaf(vL[u].y, vR[u].y, x.y);
            cl.x = fmaf(s3[u], v3[u].x, cl.x); cl.y = fmaf(s3[u], v3[u].y, cl.y);
            cr.x = fmaf(v4[u].x, s4[u], cr.x); cr.y = fmaf(v4[u].y, s4[u], cr.y);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.x + x.y + cl.x + cl.y + cr.x + cr.y;
}

int main() {
    float *out; long long *clk;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&clk, 64); cudaMemset(clk, 0, 64);
    const int iters = 2000;
    for (int stride : {43}) {
        for (int threads : {32, 64, 128, 256, 512}) {
            for (int ctas : {1, 4}) {
                long long h[3];
                cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
                cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
                cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
                probe<1><<<148 * ctas, threads, 50000>>>(out, clk, iters, stride); cudaMemcpy(&h[0], clk, 8, cudaMemcpyDeviceToHost);
                probe<2><<<148 * ctas, threads, 50000>>>(out, clk, iters, stride); cudaMemcpy(&h[1], clk, 8, cudaMemcpyDeviceToHost);
                probe<4><<<148 * ctas, threads, 50000>>>(out, clk, iters, stride); cudaMemcpy(&h[2], clk, 8, cudaMemcpyDeviceToHost);
                printf("lane stride %2d float2, %3d threads x %d CTAs/SM: clk per term  NB=1 %.1f  NB=2 %.1f  NB=4 %.1f\n", stride, threads, ctas,
                       (double)h[0] / iters, (double)h[1] / iters / 2, (double)h[2] / iters / 4);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
