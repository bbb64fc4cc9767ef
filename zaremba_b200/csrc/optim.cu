// clip_grad_norm_ + SGD over a list of tensors -- main.py:114-117.
// Pure HBM streaming.  Pass 1 reads every gradient once (sum of squares); pass 2 reads g and p and
// writes g and p (and, for the tensor-core engine, the fp16 operand images of the new weights, so the
// weights are not re-read by a separate pack pass).  Algorithmic bytes per parameter element:
// 4 (norm) + 16 (update) [+ 2..6 for fp16 images of the matrices].
// Tensors that are adjacent in memory (the Trainer's flat buffers) are coalesced into one run and
// streamed with 128-bit accesses.
#include "kernels.h"

namespace zrb {

constexpr int kNormBlocks = 148 * 8;
constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float acc, float* sh) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    float v = 0.f;
    if (threadIdx.x < 32) {
        v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
    }
    return v;
}

// sum of squares of this block's grid-stride share (blocks bx of nbx) of one run of `n` floats
__device__ __forceinline__ float block_sumsq(const float* __restrict__ g, int64_t n, int bx, int nbx, float* sh) {
    float acc = 0.f;
    const int64_t tid = (int64_t)bx * blockDim.x + threadIdx.x, stride = (int64_t)nbx * blockDim.x;
    if ((((uintptr_t)g) & 15) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const int64_t n4 = n >> 2;
        int64_t i = tid;
        for (; i + 3 * stride < n4; i += 4 * stride) {   // 4 independent 16-byte loads in flight
            float4 a = __ldcs(g4 + i), b = __ldcs(g4 + i + stride), c = __ldcs(g4 + i + 2 * stride),
                   d = __ldcs(g4 + i + 3 * stride);
            acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            acc += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
            acc += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
            acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        }
        for (; i < n4; i += stride) {
            float4 a = __ldcs(g4 + i);
            acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += stride) acc += g[j] * g[j];
    } else {
        for (int64_t j = tid; j < n; j += stride) acc += g[j] * g[j];
    }
    return block_sum(acc, sh);
}

// runs of contiguous floats (the tensor list after merging neighbours); block (x, y) takes share x of run y and owns
// partial slot y * gridDim.x + x, so one launch covers the whole list and no slot is written twice
struct Runs {
    float* p[16];
    float* g[16];
    int64_t n[16];
};
__global__ void sumsq_kernel(Runs r, float* __restrict__ partials) {
    __shared__ float sh[kThreads / 32];
    float v = block_sumsq(r.g[blockIdx.y], r.n[blockIdx.y], blockIdx.x, gridDim.x, sh);
    if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// scalars[0] = norm, scalars[1] = clip coefficient  (double accumulation of the partials)
__global__ void norm_finalize_kernel(const float* __restrict__ partials, int n, float max_norm,
                                     float* __restrict__ scalars, float* __restrict__ norm_out) {
    __shared__ double sh[32];
    double acc = 0.0;
    // up to ~15k partials: batches of 8 independent loads per thread (a rolled loop pays one L2 round trip per load)
    int i = threadIdx.x;
    for (; i + 7 * (int)blockDim.x < n; i += 8 * (int)blockDim.x) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = partials[i + k * (int)blockDim.x];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += (double)v[k];
    }
    for (; i < n; i += blockDim.x) acc += (double)partials[i];
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

// g *= coef (clip_grad_norm_ scales .grad in place); p -= lr * g (main.py:117)
template <bool WRITE_G>
__global__ void clip_sgd_update_kernel(Runs r, float lr, const float* __restrict__ scalars) {
    float* __restrict__ p = r.p[blockIdx.y];
    float* __restrict__ g = r.g[blockIdx.y];
    const int64_t n = r.n[blockIdx.y];
    const float coef = scalars[1];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (((((uintptr_t)g) | ((uintptr_t)p)) & 15) == 0) {
        float4* g4 = reinterpret_cast<float4*>(g);
        float4* p4 = reinterpret_cast<float4*>(p);
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += stride) {
            float4 gv = __ldcs(g4 + i), pv = __ldcs(p4 + i);
            gv.x *= coef; gv.y *= coef; gv.z *= coef; gv.w *= coef;
            pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
            if (WRITE_G) __stcs(g4 + i, gv);
            __stcs(p4 + i, pv);
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += stride) {
            float gv = g[j] * coef;
            if (WRITE_G) g[j] = gv;
            p[j] -= lr * gv;
        }
    } else {
        for (int64_t j = tid; j < n; j += stride) {
            float gv = g[j] * coef;
            if (WRITE_G) g[j] = gv;
            p[j] -= lr * gv;
        }
    }
}

// merge tensors that are adjacent in memory (both p and g) into runs; returns the run count and the block
// count per run (sized for the longest run, all runs' partial slots fit in kNormBlocks)
static int coalesce(const TensorList& tl, Runs* r, int* blocks_per_run) {
    int runs = 0;
    int64_t longest = 0;
    for (int t = 0; t < tl.count; ++t) {
        if (tl.n[t] == 0) continue;
        if (runs && r->p[runs - 1] + r->n[runs - 1] == tl.p[t] && r->g[runs - 1] + r->n[runs - 1] == tl.g[t]) {
            r->n[runs - 1] += tl.n[t];
        } else {
            r->p[runs] = tl.p[t]; r->g[runs] = tl.g[t]; r->n[runs] = tl.n[t];
            ++runs;
        }
    }
    for (int i = runs; i < 16; ++i) { r->p[i] = nullptr; r->g[i] = nullptr; r->n[i] = 0; }
    for (int i = 0; i < runs; ++i) longest = r->n[i] > longest ? r->n[i] : longest;
    int64_t b = (longest / 4 + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    const int cap = kNormBlocks / (runs > 0 ? runs : 1);
    *blocks_per_run = (int)(b > cap ? cap : b);
    return runs;
}

int norm_partials_base() { return kNormBlocks; }

int grad_norm(const TensorList& tl, float max_norm, float* partials, float* scalars, float* norm_out,
              cudaStream_t s, bool extra_used, int n_gemm) {
    Runs r;
    int bpr = 1;
    const int runs = coalesce(tl, &r, &bpr);
    ZRB_CUDA(cudaMemsetAsync(partials, 0, (kNormBlocks + (extra_used ? 0 : kNormExtra)) * sizeof(float), s));
    if (runs) {
        sumsq_kernel<<<dim3(bpr, runs), kThreads, 0, s>>>(r, partials);
        ZRB_KERNEL_CHECK();
    }
    // one block, 1024 threads: up to ~15k partials, a handful of independent loads per thread
    norm_finalize_kernel<<<1, 1024, 0, s>>>(partials, kNormBlocks + kNormExtra + n_gemm, max_norm, scalars, norm_out);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// update only (norm / coefficient already in `scalars`)
int sgd_apply(const TensorList& tl, float lr, const float* scalars, bool write_g, cudaStream_t s) {
    Runs r;
    int bpr = 1;
    const int runs = coalesce(tl, &r, &bpr);
    if (!runs) return ZRB_OK;
    int64_t longest = 0;   // grid.x sized for the longest run (no slot limit here); shorter runs leave blocks idle
    for (int i = 0; i < runs; ++i) longest = r.n[i] > longest ? r.n[i] : longest;
    int64_t b = (longest / 4 + kThreads - 1) / kThreads;
    const int bx = (int)(b < 1 ? 1 : (b > kNormBlocks ? kNormBlocks : b)) * 2;
    if (write_g) clip_sgd_update_kernel<true><<<dim3(bx, runs), kThreads, 0, s>>>(r, lr, scalars);
    else clip_sgd_update_kernel<false><<<dim3(bx, runs), kThreads, 0, s>>>(r, lr, scalars);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

int clip_sgd(const TensorList& tl, float lr, float max_norm, float* partials, float* scalars, float* norm_out,
             bool write_g, cudaStream_t s) {
    ZRB_TRY(grad_norm(tl, max_norm, partials, scalars, norm_out, s));
    return sgd_apply(tl, lr, scalars, write_g, s);
}

}  // namespace zrb
