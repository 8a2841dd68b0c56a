// RB_MODE_TC entry points that are not built yet report RB_ERR_UNSUPPORTED (never a silent fallback).
#include "common.cuh"
namespace rb {
int student_loss_grad_tc(int, const float*, const float*, const float*, int64_t, int, float*, float*, void*, cudaStream_t) { set_error("RB_MODE_TC student kernels are not built"); return RB_ERR_UNSUPPORTED; }
}  // namespace rb
extern "C" int rb_mode_available(int mode) { return mode == RB_MODE_FP32 || mode == RB_MODE_TC; }
extern "C" int rb_student_mode_available(int mode) { return mode == RB_MODE_FP32; }
