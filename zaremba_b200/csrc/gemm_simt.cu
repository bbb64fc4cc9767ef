// fp32 CUDA-core GEMM: the validation engine (ZRB_ENGINE_SIMT).  Exact fp32 products and
// fp32 accumulation in k order, so it tracks the fp32 oracle to rounding noise; it is the
// on-device yardstick the tcgen05 engine is compared with, not the fast path.
#include "kernels.h"

namespace zrb {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// C[M,N] = alpha * sum_k A(m,k) * B(k,n) + beta * C, A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int M, int N, int K, int64_t sam,
                                                       int64_t sak, int64_t sbk, int64_t sbn, float alpha,
                                                       float beta) {
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        // 64x16 tiles, 256 threads -> 4 elements each; pick the thread->element map so the
        // contiguous global dimension is the fastest-varying one across a warp
        for (int i = threadIdx.x; i < BM * BK; i += 256) {
            int mm, kk;
            if (sak == 1) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
            int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
        }
        for (int i = threadIdx.x; i < BN * BK; i += 256) {
            int nn, kk;
            if (sbk == 1) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? B[k * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float* c = C + (int64_t)m * N + n;
            float v = alpha * acc[i][j];
            if (beta != 0.f) v += beta * *c;
            *c = v;
        }
    }
}

int gemm_f32(const float* A, const float* B, float* C, int M, int N, int K, int transA, int transB, float alpha,
             float beta, cudaStream_t s) {
    if (M <= 0 || N <= 0) return ZRB_OK;
    int64_t sam = transA ? 1 : K, sak = transA ? M : 1;
    int64_t sbk = transB ? 1 : N, sbn = transB ? K : 1;
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    gemm_f32_kernel<<<grid, 256, 0, s>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, alpha, beta);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

}  // namespace zrb
