// ZRB_ENGINE_TC placeholder (replaced by the tcgen05 engine).
#include "engine.h"
namespace zrb {
int tc_ctx_init(zrb_ctx*) { set_error("tcgen05 engine not built yet"); return ZRB_E_INVALID; }
void tc_ctx_free(zrb_ctx*) {}
int tc_forward(zrb_ctx*, const zrb_params*, const int64_t*, const zrb_states*, const zrb_states*, float*, cudaStream_t) { return ZRB_E_INVALID; }
int tc_backward(zrb_ctx*, const zrb_params*, const float*, const zrb_params*, cudaStream_t) { return ZRB_E_INVALID; }
int tc_train_step_grads(zrb_ctx*, const zrb_params*, const zrb_params*, const int64_t*, const int64_t*, int, int, const zrb_states*, const zrb_states*, uint64_t, uint64_t, float*, cudaStream_t) { return ZRB_E_INVALID; }
int tc_update(zrb_ctx*, const zrb_params*, const TensorList&, float, float, float*, cudaStream_t) { return ZRB_E_INVALID; }
}
extern "C" int zrb_gemm_f16_tn(const void*, int64_t, const void*, int64_t, float*, int64_t, int32_t, int32_t, int32_t, float, const float*, int32_t, void*) { zrb::set_error("not built"); return ZRB_E_INVALID; }
