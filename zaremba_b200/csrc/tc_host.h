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
                float* sumsq_out = nullptr, const float* bias2 = nullptr, bool pdl = false, const __half* B2 = nullptr,
                float* C2 = nullptr, float* sumsq_out2 = nullptr, const __half* A_tiled = nullptr, int a_nt128 = 0,
                const __half* B_tiled = nullptr, int b_nt128 = 0);
// B2 / C2 (/ sumsq_out2): a second problem C2 = alpha * op(A) * op(B2)^T with the same A, shapes and pitches, computed
// by the same launch (the two weight gradients of a layer share dG as their A operand).
// pdl: launch as a programmatic dependent of the kernel enqueued just before it on `s` (which must be one of the
// persistent recurrence kernels: they release their dependents once all their CTAs are resident).  The GEMM must not
// read anything that kernel writes; it runs on the SMs the recurrence leaves idle and waits for it before completing.
// sumsq_out (plain-store calls only): gemm_f16_tc_sumsq_slots(M, N, K) floats whose sum is sum(C^2), fixed summation tree
int gemm_f16_tc_sumsq_slots(int M, int N, int K);
}  // namespace zrb
