// Shared helpers: error handling, warp-staged row I/O (coalesced [N,11] / [N,4] rows), policy parameter layout.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/reacher_b200.h"

struct rb_env {
    int64_t n = 0;
    uint64_t seed = 0;
    int device = 0;
    uint32_t offset = 0;
    float4* qv = nullptr;
    float4* tp = nullptr;
    uint4* ctr = nullptr;
    // staging for the *_host entry points
    float* d_act = nullptr; float* d_obs = nullptr; float* d_rew = nullptr; uint8_t* d_done = nullptr;
    float* d_params = nullptr;
    float* d_buf_obs = nullptr; float* d_buf_pd = nullptr; float* d_buf_rew = nullptr; uint8_t* d_buf_done = nullptr;
    int64_t buf_T = 0, buf_last_T = 0;       // allocated / last used number of time steps of the resident rollout buffer
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // D2H of finished time slabs overlaps the next slab's kernel
    cudaEvent_t slab_done[16] = {};
    int sm_count = 148;
    // In-kernel slab progress of the fused tensor-core rollout (rb_env_rollout_policy_host): every warp bumps prog_counters[slab] when it has
    // written the last step of a time slab; the warp that completes the count posts the call's epoch into prog_flags_host[slab] (page-locked
    // mapped host memory), and the host, polling it, starts the device->host copy of that slab while the SAME launch computes the next one.
    uint32_t* prog_counters = nullptr;            // device, [16], zeroed by the host call before the launch
    volatile uint32_t* prog_flags_host = nullptr; // mapped host memory, [16]
    uint32_t* prog_flags_dev = nullptr;           // device alias of prog_flags_host
    uint32_t prog_epoch = 0;
    int prog_slab_len = 0;                        // > 0 only while the host call launches its kernel
    uint64_t* done_mask_out = nullptr;            // set around a rollout launch: per-env bit mask of the steps that ended an episode (bit t = step t)
    // split-phase host rollout (rb_env_rollout_policy_host_begin / _wait): up to two calls in flight, parameters double-buffered
    float* d_params2 = nullptr;
    float* d_rew_pipe[2] = {};                    // reward staging of the two in-flight calls when the copy engine brings it to the host
    int64_t rew_pipe_rows = 0;
    cudaEvent_t pipe_kernel[2] = {};
    cudaEvent_t pipe_done[2] = {};
    uint64_t pipe_issued = 0, pipe_waited = 0;
    float* return_sum_out = nullptr;              // set around a rollout launch: per-env sum of the rewards of the launch's steps (in step order)
    float* d_return_sum = nullptr;
    uint64_t* d_done_mask = nullptr;              // device staging of that mask when the caller's buffer is not mapped
    int host_zerocopy = 1;                        // rb_env_set_host_transport: bit 0 reward, bit 1 done stored by the kernel into mapped host memory
    void* serve = nullptr;                        // resident env server of the small host-surface envs (serve.cu), NULL until first used
};

namespace rb {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

// serve.cu: the resident env server of small host-surface envs (n <= 32).  env_quiesce retires it (state back in HBM) and is called by every
// entry point that reads or writes the env state with ordinary launches; a no-op when no server is resident.
int env_quiesce(rb_env* e);
void env_serve_destroy(rb_env* e);
bool env_serve_eligible(const rb_env* e);
int env_serve_reset(rb_env* e, float* obs_host, float* pd_host);
int env_serve_step(rb_env* e, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* pd_host);

#define RB_CUDA(call)                                             \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return rb::cuda_fail(e__, #call); \
    } while (0)
#define RB_REQUIRE(cond, msg)                        \
    do {                                             \
        if (!(cond)) {                               \
            rb::set_error("%s: %s", __func__, msg);  \
            return RB_ERR_INVALID;                   \
        }                                            \
    } while (0)

// Launch state CUDA keeps PER DEVICE (function attributes, occupancy, SM count) is cached per (call site, device): the Python layer takes
// device = N everywhere, so one process may drive several GPUs.  True the first time it is called on the current device for `seen`
// (a repeated set after a race between two host threads is harmless: the cached calls are idempotent).
constexpr int RB_MAX_DEVICES = 64;
inline bool first_use_on_device(bool (&seen)[RB_MAX_DEVICES], int* device_out = nullptr) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return true; }
    if (device_out) *device_out = dev;
    if (dev < 0 || dev >= RB_MAX_DEVICES) return true;
    if (seen[dev]) return false;
    seen[dev] = true;
    return true;
}
// RAII: entry points that must select a handle's device put the caller's device back on return
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } if (prev != device) cudaSetDevice(device); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int OBS = RB_OBS_DIM;
constexpr int HID = 64;

// Warp-staged store of W floats per lane to a row-major [rows, W] global array: lanes write their row into a
// warp-private smem strip (stride W, conflict-free for odd W), then the warp streams the 32*W contiguous floats out
// with 128-bit coalesced stores.  row0 = global row of lane 0; nvalid = rows of this warp that exist (<= 32).
// Falls back to scalar stores when the destination strip is not 16-byte aligned.
template <int W>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ g, int64_t row0, int nvalid, const float* vals,
                                                float* strip, int lane) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) strip[lane * W + k] = vals[k];
    __syncwarp();
    float* dst = g + row0 * W;
    const float4* s4 = reinterpret_cast<const float4*>(strip);
    float4* d4 = reinterpret_cast<float4*>(dst);
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {      // full aligned warp (the hot case): fixed trip count, immediate offsets
        constexpr int N4 = 32 * W / 4;
#pragma unroll
        for (int u = 0; u < (N4 + 31) / 32; ++u)
            if (u * 32 + 31 < N4 || lane + u * 32 < N4) d4[lane + u * 32] = s4[lane + u * 32];
        return;
    }
    const int total = nvalid * W;
    const int n4 = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (total >> 2) : 0;
    for (int i = lane; i < n4; i += 32) d4[i] = s4[i];
    for (int i = (n4 << 2) + lane; i < total; i += 32) dst[i] = strip[i];
}

template <int W>
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ g, int64_t row0, int nvalid, float* vals, float* strip,
                                               int lane) {
    __syncwarp();
    const float* src = g + row0 * W;
    const int total = nvalid * W;
    const int n4 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (total >> 2) : 0;
    float4* s4 = reinterpret_cast<float4*>(strip);
    const float4* g4 = reinterpret_cast<const float4*>(src);
    for (int i = lane; i < n4; i += 32) s4[i] = __ldg(g4 + i);
    for (int i = (n4 << 2) + lane; i < total; i += 32) strip[i] = __ldg(src + i);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) vals[k] = lane < nvalid ? strip[lane * W + k] : 0.f;
}

// TF1 Adam update of one element (mlp_train.py:75-80; eps added to the un-corrected sqrt(v)).  Explicit roundings: the stand-alone
// kernel and the update fused into the student kernel give bit-identical parameters.
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float lr_t, float b1, float b2, float eps, float gscale) {
    const float gi = __fmul_rn(g, gscale);
    m = __fmaf_rn(b1, m, __fmul_rn(__fsub_rn(1.f, b1), gi));
    v = __fmaf_rn(b2, v, __fmul_rn(__fsub_rn(1.f, b2), __fmul_rn(gi, gi)));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
}

// ---- policy parameters (MlpPolicy 11-64-64-nout) ------------------------------------------------------------
// global flat layout (include/reacher_b200.h): ob_mean[11] ob_std[11] W1[11][64] b1[64] W2[64][64] b2[64] W3[64][nout]
// b3[nout] logstd[2]
struct PolicyOffsets {
    int mu, sd, W1, b1, W2, b2, W3, b3, logstd, total;
};
__host__ __device__ inline PolicyOffsets policy_offsets(int nout) {
    PolicyOffsets o;
    o.mu = 0; o.sd = 11; o.W1 = 22; o.b1 = o.W1 + 11 * HID; o.W2 = o.b1 + HID; o.b2 = o.W2 + HID * HID; o.W3 = o.b2 + HID;
    o.b3 = o.W3 + HID * nout; o.logstd = o.b3 + nout; o.total = o.logstd + 2;
    return o;
}

}  // namespace rb
