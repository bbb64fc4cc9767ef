// Launchers used only by the tcgen05 engine.
#pragma once
#include "kernels.h"
#include "tc_host.h"

namespace zrb {

constexpr float kGradScale = 1024.f;   // fp16 gradient images hold kGradScale * value (exact power of two)

int convert_pad_f16(const float* src, int64_t ld_src, __half* dst, int64_t ld_dst, int rows, int cols, float scale,
                    cudaStream_t s);
int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s);
int colsum_h(const __half* A, int64_t ld, float* out, int N, int M, float inv_scale, cudaStream_t s);

// cell pointwise with fp16 side outputs (tc_cell.cu)
int lstm_cell_fwd_tc(float* pre, const float* c_prev, float* c_out, float* h_raw, __half* h_raw_h, __half* y_h,
                     int64_t ld_h, int B, int H, int64_t elem_off, int64_t n_total, MaskSrc m, cudaStream_t s);
int lstm_cell_bwd_tc(const float* dy_post, const float* dh_rec, float* dc, const float* gates, const float* c_t,
                     const float* c_prev, float* dG, __half* dG_h, int64_t ld_g, int B, int H, int64_t elem_off,
                     int64_t n_total, MaskSrc m, cudaStream_t s);

}  // namespace zrb
