// Host-side helpers shared by the tcgen05 kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace zrb {
int tc_num_sms();
int tc_make_tmap_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                     uint32_t box_outer, int swizzle128);
// C[M,N] fp32 = alpha * op(A) * op(B)^T (+bias) (+C); *_mn = operand stored with the M/N index contiguous
int gemm_f16_tc(const __half* A, int64_t lda, int a_mn, const __half* B, int64_t ldb, int b_mn, float* C, int64_t ldc,
                int M, int N, int K, float alpha, const float* bias, int accumulate, cudaStream_t s,
                float* sumsq_out = nullptr, const float* bias2 = nullptr);
// sumsq_out (plain-store calls only): gemm_f16_tc_sumsq_slots(M, N, K) floats whose sum is sum(C^2), fixed summation tree
int gemm_f16_tc_sumsq_slots(int M, int N, int K);
}  // namespace zrb
