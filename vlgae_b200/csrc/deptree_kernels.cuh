// deptree_kernels.cuh -- helpers that map the arc-factored (MBR) chart onto the DMV kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vlgae {

// attach2 [B][N][N][2] = (arc, arc), dec [B][N][8] = 0;  marg [B][N][N] = g2[..., 0] + g2[..., 1]
cudaError_t launch_deptree_expand(const float *arc, int B, int N, float *attach2, float *dec, cudaStream_t st);
cudaError_t launch_deptree_collapse(const float *g2, int B, int N, float *marg, cudaStream_t st);

}  // namespace vlgae
