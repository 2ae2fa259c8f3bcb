// deptree_kernels.cuh -- internal interface of the arc-factored (MBR) chart kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vlgae {

struct DepTreeArgs {
    const float *arc;        // [B][N][N] (head, child)
    const int64_t *lengths;  // [B]
    int B, N;
    float fill, mask_zero;
    float *out;      // [B]
    float *marg;     // [B][N][N] or null
    int64_t *heads;  // [B][N] or null
    void *workspace;
    size_t ws_stride;
};

size_t deptree_ws_stride(int N);
cudaError_t launch_deptree(const DepTreeArgs &a, int semiring, cudaStream_t st);

}  // namespace vlgae
