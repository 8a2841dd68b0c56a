// CUDA-core fp32 evaluation of the baselines MlpPolicy (11 -> 64 tanh -> 64 tanh -> nout) -- RB_MODE_FP32.
// Replaces sess.run((pi.pd.mean, pi.pd.flat)) at /root/reference src/distilation/mlp_train.py:123-125,165-167
// (network built by teacher.py:12-16).  One thread per sample, activations in registers, weights broadcast from
// shared memory with 128-bit loads.  This is the strict-parity path; the tensor-core path lives in policy_tc.cuh.
#pragma once
#include "common.cuh"

namespace rb {

// shared-memory image of the policy: W3/b3 padded to 4 outputs (zero columns when nout == 2)
struct __align__(16) PolicySmem {
    float W1[11][HID];
    float b1[HID];
    float W2[HID][HID];
    float b2[HID];
    float W3[HID][4];
    float b3[4];
    float mu[12];
    float inv_sd[12];
    float logstd[4];
};

__device__ inline void policy_load_smem(PolicySmem& S, const float* __restrict__ p, int nout) {
    const PolicyOffsets o = policy_offsets(nout);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 11 * HID; i += nt) (&S.W1[0][0])[i] = __ldg(p + o.W1 + i);
    for (int i = tid; i < HID * HID; i += nt) (&S.W2[0][0])[i] = __ldg(p + o.W2 + i);
    for (int i = tid; i < HID; i += nt) {
        S.b1[i] = __ldg(p + o.b1 + i);
        S.b2[i] = __ldg(p + o.b2 + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) S.W3[i][j] = j < nout ? __ldg(p + o.W3 + i * nout + j) : 0.f;
    }
    if (tid < 4) {
        S.b3[tid] = tid < nout ? __ldg(p + o.b3 + tid) : 0.f;
        S.logstd[tid] = tid < 2 ? __ldg(p + o.logstd + tid) : 0.f;
    }
    if (tid < 12) {
        S.mu[tid] = tid < 11 ? __ldg(p + o.mu + tid) : 0.f;
        S.inv_sd[tid] = tid < 11 ? 1.0f / __ldg(p + o.sd + tid) : 0.f;
    }
}

// pd[4] = (mean0, mean1, logstd0, logstd1) for NOUT == 2, the raw four outputs for NOUT == 4
template <int NOUT> __device__ __forceinline__ void policy_fwd_simt(const PolicySmem& S, const float* ob, float* pd) {
    float z[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float v = (ob[k] - S.mu[k]) * S.inv_sd[k];
        z[k] = fminf(5.f, fmaxf(-5.f, v));
    }
    float h1[HID];
#pragma unroll
    for (int j4 = 0; j4 < HID / 4; ++j4) {
        const float4 b = *reinterpret_cast<const float4*>(&S.b1[4 * j4]);
        h1[4 * j4] = b.x; h1[4 * j4 + 1] = b.y; h1[4 * j4 + 2] = b.z; h1[4 * j4 + 3] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) {
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(&S.W1[k][4 * j4]);
            h1[4 * j4] = fmaf(z[k], w.x, h1[4 * j4]);
            h1[4 * j4 + 1] = fmaf(z[k], w.y, h1[4 * j4 + 1]);
            h1[4 * j4 + 2] = fmaf(z[k], w.z, h1[4 * j4 + 2]);
            h1[4 * j4 + 3] = fmaf(z[k], w.w, h1[4 * j4 + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < HID; ++j) h1[j] = tanhf(h1[j]);
    float h2[HID];
#pragma unroll
    for (int j4 = 0; j4 < HID / 4; ++j4) {
        const float4 b = *reinterpret_cast<const float4*>(&S.b2[4 * j4]);
        h2[4 * j4] = b.x; h2[4 * j4 + 1] = b.y; h2[4 * j4 + 2] = b.z; h2[4 * j4 + 3] = b.w;
    }
#pragma unroll
    for (int k = 0; k < HID; ++k) {
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(&S.W2[k][4 * j4]);
            h2[4 * j4] = fmaf(h1[k], w.x, h2[4 * j4]);
            h2[4 * j4 + 1] = fmaf(h1[k], w.y, h2[4 * j4 + 1]);
            h2[4 * j4 + 2] = fmaf(h1[k], w.z, h2[4 * j4 + 2]);
            h2[4 * j4 + 3] = fmaf(h1[k], w.w, h2[4 * j4 + 3]);
        }
    }
    float o0 = S.b3[0], o1 = S.b3[1], o2 = S.b3[2], o3 = S.b3[3];
#pragma unroll
    for (int k = 0; k < HID; ++k) {
        const float h = tanhf(h2[k]);
        const float4 w = *reinterpret_cast<const float4*>(&S.W3[k][0]);
        o0 = fmaf(h, w.x, o0);
        o1 = fmaf(h, w.y, o1);
        if (NOUT == 4) { o2 = fmaf(h, w.z, o2); o3 = fmaf(h, w.w, o3); }
    }
    pd[0] = o0; pd[1] = o1;
    if (NOUT == 4) { pd[2] = o2; pd[3] = o3; } else { pd[2] = S.logstd[0]; pd[3] = S.logstd[1]; }
}

}  // namespace rb
