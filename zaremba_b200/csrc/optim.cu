// clip_grad_norm_ + SGD over a list of tensors -- main.py:114-117.
// Pure HBM streaming: pass 1 reads every gradient once (sum of squares), pass 2 reads
// g and p and writes g and p.  Algorithmic bytes per parameter element: 4 (norm) + 16.
#include "kernels.h"

namespace zrb {

constexpr int kNormBlocks = 592;  // 4 x 148 SMs
constexpr int kThreads = 256;

// flat virtual index space over all tensors; each block walks a contiguous slice of it
__global__ void sumsq_kernel(TensorList tl, int64_t total, float* __restrict__ partials) {
    __shared__ float sh[kThreads / 32];
    int64_t per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + 3) & ~(int64_t)3;
    int64_t lo = per * blockIdx.x, hi = lo + per < total ? lo + per : total;
    float acc = 0.f;
    int64_t base = 0;
    for (int t = 0; t < tl.count; ++t) {
        int64_t n = tl.n[t];
        int64_t a = lo > base ? lo : base, b = hi < base + n ? hi : base + n;
        if (a < b) {
            const float* g = tl.g[t] - base;
            for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) {
                float v = g[i];
                acc += v * v;
            }
        }
        base += n;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

// scalars[0] = norm, scalars[1] = clip coefficient  (double accumulation of the partials)
__global__ void norm_finalize_kernel(const float* __restrict__ partials, int n, float max_norm,
                                     float* __restrict__ scalars, float* __restrict__ norm_out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partials[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        float norm = (float)sqrt(t);
        float coef = max_norm / (norm + 1e-6f);   // torch.nn.utils.clip_grad_norm_
        if (coef > 1.f) coef = 1.f;
        scalars[0] = norm;
        scalars[1] = coef;
        if (norm_out) *norm_out = norm;
    }
}

__global__ void clip_sgd_update_kernel(TensorList tl, int64_t total, float lr, const float* __restrict__ scalars) {
    float coef = scalars[1];
    int64_t per = (total + gridDim.x - 1) / gridDim.x;
    int64_t lo = per * blockIdx.x, hi = lo + per < total ? lo + per : total;
    int64_t base = 0;
    for (int t = 0; t < tl.count; ++t) {
        int64_t n = tl.n[t];
        int64_t a = lo > base ? lo : base, b = hi < base + n ? hi : base + n;
        if (a < b) {
            float* g = tl.g[t] - base;
            float* p = tl.p[t] - base;
            for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) {
                float gv = g[i] * coef;     // clip_grad_norm_ scales .grad in place
                g[i] = gv;
                p[i] -= lr * gv;            // main.py:117
            }
        }
        base += n;
    }
}

int clip_sgd(const TensorList& tl, float lr, float max_norm, float* partials, float* scalars, float* norm_out,
             cudaStream_t s) {
    int64_t total = 0;
    for (int t = 0; t < tl.count; ++t) total += tl.n[t];
    if (!total) return ZRB_OK;
    sumsq_kernel<<<kNormBlocks, kThreads, 0, s>>>(tl, total, partials);
    ZRB_KERNEL_CHECK();
    norm_finalize_kernel<<<1, 256, 0, s>>>(partials, kNormBlocks, max_norm, scalars, norm_out);
    ZRB_KERNEL_CHECK();
    clip_sgd_update_kernel<<<kNormBlocks * 2, kThreads, 0, s>>>(tl, total, lr, scalars);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

}  // namespace zrb
