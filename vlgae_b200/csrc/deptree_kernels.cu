// deptree_kernels.cu -- arc-factored projective dependency CRF (Eisner) for MBR decoding.
//
// Replaces  DepTree._dp / _check_potentials   /root/reference/src/model/torch_struct/deptree.py:25-76,146-162
//           DependencyCRF.partition/max/argmax/marginals   distributions.py:269-299 (+ helpers.py:118-154)
// Only reached with `mbr_decoding: true` (config/model/vlgae.yaml:85, default false), so this kernel is written for
// correctness and simplicity: one CTA per sentence, one thread per span, charts in a global workspace (L2-resident),
// two barriers per width; the reverse sweep is explicit (no autograd) and conflict-free by the same ownership argument
// as the DMV kernel.  Runs every sentence at its own N_b = len + 1.
#include <cuda_runtime.h>
#include <stdint.h>

#include "deptree_kernels.cuh"

namespace vlgae {
namespace {

constexpr float NEG_BIG = -3.0e38f;

struct Red {  // log-sum-exp or first-max over a strided list of terms
    float m, s;
    int a;
};

template <bool MAX>
__device__ __forceinline__ float reduce_terms(const float *x, int xs, const float *y, int ys, int n, int *arg) {
    // terms t_r = x[r * xs] + y[r * ys], r = 0..n-1
    float m = NEG_BIG;
    int a = 0;
    for (int r = 0; r < n; ++r) {
        const float t = x[(size_t)r * xs] + y[(size_t)r * ys];
        if (t > m) { m = t; a = r; }  // strict: first maximum (torch.max tie rule)
    }
    if (MAX) { *arg = a; return m; }
    float s = 0.f;
    for (int r = 0; r < n; ++r) s += __expf(x[(size_t)r * xs] + y[(size_t)r * ys] - m);
    return m + __logf(s);
}

template <bool MAX>
__global__ void __launch_bounds__(128) deptree_kernel(DepTreeArgs p) {
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int N = p.N;
    int len = (int)p.lengths[b];
    len = len < 0 ? 0 : (len > N - 1 ? N - 1 : len);
    const int Nb = len + 1;
    const size_t n2 = (size_t)Nb * Nb;
    float *C = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(p.workspace) + (size_t)b * p.ws_stride);
    float *I = C + n2, *X = I + n2, *gC = X + n2, *gI = gC + n2;
    int *bpX = reinterpret_cast<int *>(gI + n2), *bpC = bpX + n2;
    const float *arc = p.arc + (size_t)b * N * N;
#define AT(Q, r, c) Q[(size_t)(r) * Nb + (c)]
    const bool want_back = p.marg != nullptr || p.heads != nullptr;
    for (size_t t = tid; t < n2; t += NT) { C[t] = p.fill; I[t] = p.fill; gC[t] = 0.f; gI[t] = 0.f; }
    __syncthreads();
    for (int i = tid; i < Nb; i += NT) AT(C, i, i) = 0.f;  // deptree.py:44
    __syncthreads();
    for (int w = 1; w < Nb; ++w) {
        for (int i = tid; i + w < Nb; i += NT) {
            const int j = i + w;
            int arg = 0;
            // X = (+)_{r=i..j-1} C[i][r] + C[j][r+1]          (deptree.py:53-55)
            const float x = reduce_terms<MAX>(&AT(C, i, i), 1, &AT(C, j, i + 1), 1, w, &arg);
            AT(X, i, j) = x; AT(bpX, i, j) = arg;
            AT(I, j, i) = x + arc[(size_t)j * N + i];  // deptree.py:59
            AT(I, i, j) = x + arc[(size_t)i * N + j];  // deptree.py:63
        }
        __syncthreads();
        for (int i = tid; i + w < Nb; i += NT) {
            const int j = i + w;
            int arg = 0;
            // C[j][i] = (+)_{r=i..j-1} C[r][i] + I[j][r]     (deptree.py:66-67)
            AT(C, j, i) = reduce_terms<MAX>(&AT(C, i, i), Nb, &AT(I, j, i), 1, w, &arg);
            AT(bpC, j, i) = arg;
            // C[i][j] = (+)_{r=i+1..j} I[i][r] + C[r][j]     (deptree.py:69-70)
            float c = reduce_terms<MAX>(&AT(I, i, i + 1), 1, &AT(C, i + 1, j), Nb, w, &arg);
            AT(bpC, i, j) = arg;
            if (i == 0 && w != len) c = p.mask_zero;           // deptree.py:72-73
            AT(C, i, j) = c;
        }
        __syncthreads();
    }
    if (tid == 0) p.out[b] = AT(C, 0, len);
    if (!want_back) return;

    if (tid == 0) AT(gC, 0, len) = 1.f;
    __syncthreads();
    for (int w = Nb - 1; w >= 1; --w) {
        for (int i = tid; i + w < Nb; i += NT) {
            const int j = i + w;
            float g = AT(gC, i, j);
            if (i == 0 && w != len) g = 0.f;
            if (g != 0.f) {
                const float out = AT(C, i, j);
                for (int r = i + 1; r <= j; ++r) {
                    const float pr = MAX ? (r - i - 1 == AT(bpC, i, j) ? g : 0.f) : g * __expf(AT(I, i, r) + AT(C, r, j) - out);
                    AT(gI, i, r) += pr; AT(gC, r, j) += pr;
                }
            }
            g = AT(gC, j, i);
            if (g != 0.f) {
                const float out = AT(C, j, i);
                for (int r = i; r < j; ++r) {
                    const float pr = MAX ? (r - i == AT(bpC, j, i) ? g : 0.f) : g * __expf(AT(C, r, i) + AT(I, j, r) - out);
                    AT(gC, r, i) += pr; AT(gI, j, r) += pr;
                }
            }
        }
        __syncthreads();
        for (int i = tid; i + w < Nb; i += NT) {
            const int j = i + w;
            const float gx = AT(gI, i, j) + AT(gI, j, i);
            if (gx != 0.f) {
                const float out = AT(X, i, j);
                for (int r = i; r < j; ++r) {
                    const float pr = MAX ? (r - i == AT(bpX, i, j) ? gx : 0.f) : gx * __expf(AT(C, i, r) + AT(C, j, r + 1) - out);
                    AT(gC, i, r) += pr; AT(gC, j, r + 1) += pr;
                }
            }
        }
        __syncthreads();
    }
    if (p.marg) {
        float *mg = p.marg + (size_t)b * N * N;
        for (int t = tid; t < N * N; t += NT) {
            const int h = t / N, c = t - h * N;
            mg[t] = (h < Nb && c < Nb && h != c) ? AT(gI, h, c) : 0.f;
        }
    }
    if (p.heads) {
        int64_t *hd = p.heads + (size_t)b * N;
        for (int c = tid; c < N; c += NT) {
            int64_t h = 0;
            if (c >= 1 && c < Nb)
                for (int k = 0; k < Nb; ++k)
                    if (k != c && AT(gI, k, c) != 0.f) h = k;
            hd[c] = h;
        }
    }
#undef AT
}

}  // namespace

size_t deptree_ws_stride(int N) { return (((size_t)N * N * 7 * 4) + 255) & ~(size_t)255; }

cudaError_t launch_deptree(const DepTreeArgs &a, int semiring, cudaStream_t st) {
    if (a.B == 0) return cudaSuccess;
    if (semiring == 1) deptree_kernel<true><<<a.B, 128, 0, st>>>(a);
    else deptree_kernel<false><<<a.B, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace vlgae
