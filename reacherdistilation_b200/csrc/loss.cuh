// KL between diagonal Gaussians given as pdflat rows (mean0, mean1, logstd0, logstd1), summed over the two dimensions, and its
// gradient with respect to the student row: /root/reference src/distilation/loss.py:8-13 (KL(student || teacher)) and
// backup/student_rollout.py:639-640 (KL(teacher || student)).
#pragma once
#include "../../include/reacher_b200.h"

namespace rb {

__device__ __forceinline__ float kl_row(const float4 sv, const float4 tv, int loss_kind, float4& d) {
    const float vs0 = expf(2.f * sv.z), vs1 = expf(2.f * sv.w), vt0 = expf(2.f * tv.z), vt1 = expf(2.f * tv.w);
    const float e0 = sv.x - tv.x, e1 = sv.y - tv.y;
    if (loss_kind == RB_LOSS_KL_ST) {
        d = make_float4(e0 / vt0, e1 / vt1, vs0 / vt0 - 1.f, vs1 / vt1 - 1.f);
        return (tv.z - sv.z + (vs0 + e0 * e0) / (2.f * vt0) - 0.5f) + (tv.w - sv.w + (vs1 + e1 * e1) / (2.f * vt1) - 0.5f);
    }
    d = make_float4(e0 / vs0, e1 / vs1, 1.f - (vt0 + e0 * e0) / vs0, 1.f - (vt1 + e1 * e1) / vs1);
    return (sv.z - tv.z + (vt0 + e0 * e0) / (2.f * vs0) - 0.5f) + (sv.w - tv.w + (vt1 + e1 * e1) / (2.f * vs1) - 0.5f);
}

// any loss kind a pdflat student can be trained with: the two KL directions, the squared error on the whole pdflat row (RB_LOSS_MSE)
// or on its mean half only -- the ACTIONS, north_star (2) "MSE/KL on actions" (RB_LOSS_MSE_ACTION: the logstd outputs get no gradient)
__device__ __forceinline__ float pd_loss_row(const float4 sv, const float4 tv, int loss_kind, float4& d) {
    if (loss_kind == RB_LOSS_MSE || loss_kind == RB_LOSS_MSE_ACTION) {
        const float e0 = sv.x - tv.x, e1 = sv.y - tv.y;
        const float e2 = loss_kind == RB_LOSS_MSE ? sv.z - tv.z : 0.f, e3 = loss_kind == RB_LOSS_MSE ? sv.w - tv.w : 0.f;
        d = make_float4(2.f * e0, 2.f * e1, 2.f * e2, 2.f * e3);
        return (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
    }
    return kl_row(sv, tv, loss_kind, d);
}
__host__ __device__ inline bool pd_loss_kind_ok(int k) { return k == RB_LOSS_KL_ST || k == RB_LOSS_KL_TS || k == RB_LOSS_MSE || k == RB_LOSS_MSE_ACTION; }

}  // namespace rb
