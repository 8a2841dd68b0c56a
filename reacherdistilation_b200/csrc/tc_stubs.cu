// Placeholder until the tcgen05 paths land: RB_MODE_TC entry points report RB_ERR_UNSUPPORTED (never a silent fallback).
#include "common.cuh"
namespace rb {
int policy_fwd_tc(const float*, int, const float*, int64_t, float*, cudaStream_t) { set_error("RB_MODE_TC policy_fwd not built"); return RB_ERR_UNSUPPORTED; }
int rollout_policy_tc(rb_env*, const float*, int, int, float*, float*, float*, uint8_t*, cudaStream_t) { set_error("RB_MODE_TC rollout not built"); return RB_ERR_UNSUPPORTED; }
int student_loss_grad_tc(int, const float*, const float*, const float*, int64_t, int, float*, float*, void*, cudaStream_t) { set_error("RB_MODE_TC student not built"); return RB_ERR_UNSUPPORTED; }
}  // namespace rb
extern "C" int rb_mode_available(int mode) { return mode == RB_MODE_FP32; }
