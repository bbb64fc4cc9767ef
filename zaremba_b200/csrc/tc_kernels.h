// Launchers used only by the tcgen05 engine.
#pragma once
#include "kernels.h"
#include "tc_host.h"

namespace zrb {

constexpr float kGradScale = 1024.f;   // fp16 gradient images hold kGradScale * value (exact power of two)

int convert_pad_f16(const float* src, int64_t ld_src, __half* dst, int64_t ld_dst, int rows, int cols, float scale,
                    cudaStream_t s);
int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s);
int colsum_h(const __half* A, int64_t ld, float* out, float* out2, int N, int M, float inv_scale, cudaStream_t s);

// cell pointwise with fp16 side outputs (tc_cell.cu)
int lstm_cell_fwd_tc(float* pre, const float* c_prev, float* c_out, float* h_raw, __half* h_raw_h, __half* y_h,
                     int64_t ld_h, int B, int H, int64_t elem_off, int64_t n_total, MaskSrc m, cudaStream_t s);
int lstm_cell_bwd_tc(const float* dy_post, const float* dh_rec, float* dc, const float* gates, const float* c_t,
                     const float* c_prev, float* dG, __half* dG_h, int64_t ld_g, int B, int H, int64_t elem_off,
                     int64_t n_total, MaskSrc m, cudaStream_t s);

// ---- persistent recurrence (lstm_rec_fwd.cu / lstm_rec_bwd.cu) ---------------------------------------
struct RecPlan {
    int ok;      // shape fits the persistent kernel (else the per-timestep path is used)
    int U;       // hidden units per CTA
    int G;       // 8-row groups of the weight slice (ceil(4U/8))
    int GB;      // 8-row groups of the batch operand (ceil(B/8))
    int Kc;      // 8-element K chunks (ceil16(H)/8)
    int nCTA;
    int smem;
};
size_t rec_smem_bytes(int Kc, int G, int GB);
int rec_fwd_plan(int H, int B, RecPlan* plan);
int pack_whh_fwd(const float* W, __half* img, int H, const RecPlan& p, cudaStream_t s);
int pack_h_image(const float* h, __half* img, int B, int H, const RecPlan& p, cudaStream_t s);
int lstm_rec_fwd(const RecPlan& p, const __half* w_img, __half* h_img, float* gates, const float* c0, float* cst,
                 float* h_last, float* c_last, __half* hprev_h, __half* y_h, unsigned int* counter, int T, int B, int H,
                 int Hp, MaskSrc m, cudaStream_t s, long long* trace = nullptr);
// SGD update of one matrix fused with its fp16 image rebuild (optim_tc.cu)
int update_pack(float* p, float* g, int rows, int cols, float lr, const float* scalars, __half* row_img, int64_t ld,
                __half* fwd_img, const RecPlan* fp, __half* bwd_img, const RecPlan* bp, bool write_g, cudaStream_t s);
int rec_bwd_plan(int H, int B, RecPlan* plan);   // U = units per CTA, nCTA = 4 * clusters
int pack_whh_bwd(const float* W, __half* img, int H, const RecPlan& p, cudaStream_t s);
int lstm_rec_bwd(const RecPlan& p, const __half* w_img, __half* g_img, const float* dy, const float* gates,
                 const float* cst, const float* c0, __half* dG_h, unsigned int* counter, int T, int B, int H, int G4p,
                 MaskSrc m, cudaStream_t s, long long* trace = nullptr);

}  // namespace zrb
