// tools/probes/store_pattern.cu -- how fast can 148 persistent CTAs write the alignment output [B,A,Q,ldv] when each
// CTA owns a W-float wide column strip of one image (the tiling the tcgen05 kernel has, W = 128), versus wider strips.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/store_pattern tools/probes/store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC>
__global__ void __launch_bounds__(256) strip_writer(float *out, int A, int B, int Q, int ldv, int W, int nstrip, int warps_used) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps_used) return;
    const int n_items = A * nstrip;
    const int per_row = W / (32 * VEC);  // warp-stores per row of the strip
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int a = item / nstrip, s = item - a * nstrip;
        for (int b = 0; b < B; ++b) {
            float *base = out + ((size_t)(b * A + a) * Q) * ldv + (size_t)s * W;
            for (int u = warp; u < Q * per_row; u += warps_used) {
                const int q = u / per_row, c = u - q * per_row;
                const int col = (c * 32 + lane) * VEC;
                if (s * W + col < ldv) {
                    float *o = base + (size_t)q * ldv + col;
                    if (VEC == 1) __stcs(o, 1.0f);
                    else __stcs(reinterpret_cast<float4 *>(o), make_float4(1.f, 1.f, 1.f, 1.f));
                }
            }
        }
    }
}

int main(int argc, char **argv) {
    const int A = 128, B = 128, Q = 82, ldv = 1376;
    const size_t n = (size_t)A * B * Q * ldv;
    float *out;
    cudaMalloc(&out, n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int Ws[] = {128, 256, 512, 1408};
    for (int vec = 1; vec <= 4; vec += 3)
        for (int wi = 0; wi < 4; ++wi)
            for (int warps = 8; warps <= 8; warps += 4)
                for (int grid = 148; grid <= 592; grid *= 2) {
                    const int W = Ws[wi];
                    const int nstrip = (ldv + W - 1) / W;
                    float ms = 0;
                    for (int it = 0; it < 3; ++it) {
                        cudaEventRecord(e0);
                        if (vec == 1) strip_writer<1><<<grid, 256>>>(out, A, B, Q, ldv, W, nstrip, warps);
                        else strip_writer<4><<<grid, 256>>>(out, A, B, Q, ldv, W, nstrip, warps);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        cudaEventElapsedTime(&ms, e0, e1);
                    }
                    printf("vec=%d W=%4d warps=%d grid=%3d: %.3f ms  %.0f GB/s\n", vec, W, warps, grid, ms, n * 4 / ms / 1e6);
                }
    cudaEventRecord(e0);
    cudaMemsetAsync(out, 0, n * 4);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
