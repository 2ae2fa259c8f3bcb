// deptree_kernels.cu -- arc-factored projective dependency CRF (MBR decoding) on the DMV kernels.
//
// Replaces  DepTree._dp / DependencyCRF   /root/reference/src/model/torch_struct/deptree.py:25-76,146-162
// (+ the autograd marginals / argmax of helpers.py:118-154).
#include <cuda_runtime.h>
#include <stdint.h>

#include "deptree_kernels.cuh"

namespace vlgae {

// ---------------------------------------------------------------------------------------------
// The arc-factored chart is the DMV chart without valence (the same steps in the same split order, the same single-root
// mask: deptree.py:52-73 vs dmv.py:50-63).  With dec == 0 and attach[h][c][HAS] = attach[h][c][NO] = arc[h][c] both
// valences of every DMV item carry bit for bit the value of the corresponding DepTree item (x + 0.0f == x), so the DMV
// kernels compute it; the two valence slots of the gradient / indicator add up to the arc marginal / the 0-1 indicator.
// ---------------------------------------------------------------------------------------------
__global__ void deptree_expand_kernel(const float *arc, size_t n_arc, float *attach2, float *dec, size_t n_dec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_arc; t += stride) {
        const float x = arc[t];
        reinterpret_cast<float2 *>(attach2)[t] = make_float2(x, x);
    }
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_dec; t += stride) dec[t] = 0.f;
}
__global__ void deptree_collapse_kernel(const float *g2, size_t n_arc, float *marg) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n_arc; t += stride) {
        const float2 v = reinterpret_cast<const float2 *>(g2)[t];
        marg[t] = v.x + v.y;
    }
}

cudaError_t launch_deptree_expand(const float *arc, int B, int N, float *attach2, float *dec, cudaStream_t st) {
    const size_t n_arc = (size_t)B * N * N, n_dec = (size_t)B * N * 8;
    int grid = (int)((n_arc + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    deptree_expand_kernel<<<grid, 256, 0, st>>>(arc, n_arc, attach2, dec, n_dec);
    return cudaGetLastError();
}
cudaError_t launch_deptree_collapse(const float *g2, int B, int N, float *marg, cudaStream_t st) {
    const size_t n_arc = (size_t)B * N * N;
    int grid = (int)((n_arc + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    deptree_collapse_kernel<<<grid, 256, 0, st>>>(g2, n_arc, marg);
    return cudaGetLastError();
}

}  // namespace vlgae
