// Student input row of the MLP student: x = concat(dropout(ob, keep_prob), prev_pdflat, prev_rew)
// (/root/reference src/distilation/mlp_train.py:50-52; tf.nn.dropout: x / kp * floor(kp + u)); u from Philox keyed
// (seed, global sample id, iteration, draw) so the mask is bit-identical in the oracle (oracle/nn_np.py student_input).
#pragma once
#include "philox.cuh"

namespace rb {

__device__ __forceinline__ void mlp_input_row(const float* ob11, float keep_prob, uint32_t k0, uint32_t k1, uint32_t sample_id, uint32_t iteration,
                                              float4 pp, float pr, float4* out4) {
    float ob[12];
#pragma unroll
    for (int k = 0; k < 11; ++k) ob[k] = ob11[k];
    ob[11] = 0.f;
    if (keep_prob < 1.f) {
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            const uint4 r = philox4x32_10(sample_id, iteration, (uint32_t)blk, STREAM_DROPOUT, k0, k1);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float u = (float)(rr[c] >> 8) * 5.9604644775390625e-08f;
                ob[4 * blk + c] = __fdiv_rn(ob[4 * blk + c], keep_prob) * floorf(keep_prob + u);
            }
        }
    }
    out4[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
    out4[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
    out4[2] = make_float4(ob[8], ob[9], ob[10], pp.x);
    out4[3] = make_float4(pp.y, pp.z, pp.w, pr);
}

}  // namespace rb
