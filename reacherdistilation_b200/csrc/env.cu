// Lock-step Reacher-v2 environments on the device: reset / step / observe / fused multi-step rollouts.
// Replaces the gym env calls of the reference loops (/root/reference src/distilation/mlp_train.py:21,112,135,138,196,200;
// lstm_train.py:21,111,133,136,192,196) -- see include/reacher_b200.h for the per-entry-point mapping.
//
// HBM layout (one env per thread, 128-bit coalesced state I/O):
//   qv  float4[N] = (q0, q1, v0, v1)   tp  float4[N] = (tx, ty, px, py)   ctr uint4[N] = (step, episode, low parts of q0 / q1: physics.cuh)
// Observation rows ([N,11] fp32, 44 B) are staged through a warp-private shared-memory strip so that the global
// stores are 128-bit and contiguous.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "physics.cuh"
#include "policy_simt.cuh"

namespace rb {

constexpr int ENV_BLOCK = 128;

__global__ void __launch_bounds__(ENV_BLOCK) k_reset(int64_t n, float4* qv, float4* tp, uint4* ctr, float* obs, uint32_t k0,
                                                     uint32_t k1, uint32_t offset) {
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * ENV_BLOCK + threadIdx.x;
    const int64_t row0 = i - lane;
    if (row0 >= n) return;
    const int nvalid = (int)min((int64_t)32, n - row0);
    float ob[OBS];
    if (i < n) {
        EnvState e;
        e.episode = 0u;
        reset_env(e, k0, k1, offset + (uint32_t)i);
        store_state(qv, tp, ctr, i, e);
        observe(e, ob);
    }
    if (obs) warp_store_rows<OBS>(obs, row0, nvalid, ob, strips[warp], lane);
}

__global__ void __launch_bounds__(ENV_BLOCK) k_observe(int64_t n, const float4* qv, const float4* tp, const uint4* ctr, float* obs) {
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * ENV_BLOCK + threadIdx.x;
    const int64_t row0 = i - lane;
    if (row0 >= n) return;
    const int nvalid = (int)min((int64_t)32, n - row0);
    float ob[OBS];
    if (i < n) { const EnvState e = load_state(qv, tp, ctr, i); observe(e, ob); }
    warp_store_rows<OBS>(obs, row0, nvalid, ob, strips[warp], lane);
}

// env.step: the single-step API kernel (HBM-bound: 113 algorithmic bytes per env-step)
__global__ void __launch_bounds__(ENV_BLOCK) k_step(int64_t n, float4* __restrict__ qv, float4* __restrict__ tp, uint4* __restrict__ ctr,
                                                    const float2* __restrict__ act, float* __restrict__ obs, float* __restrict__ rew,
                                                    uint8_t* __restrict__ done, uint32_t k0, uint32_t k1, uint32_t offset) {
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * ENV_BLOCK + threadIdx.x;
    const int64_t row0 = i - lane;
    if (row0 >= n) return;
    const int nvalid = (int)min((int64_t)32, n - row0);
    float ob[OBS];
    if (i < n) {
        EnvState e = load_state(qv, tp, ctr, i);
        const float2 a = __ldg(act + i);
        bool d;
        const float r = step_env(e, a.x, a.y, k0, k1, offset + (uint32_t)i, d);
        store_state(qv, tp, ctr, i, e);
        if (rew) rew[i] = r;
        if (done) done[i] = d ? 1 : 0;
        observe(e, ob);
    }
    if (obs) warp_store_rows<OBS>(obs, row0, nvalid, ob, strips[warp], lane);
}

__global__ void k_get_state(int64_t n, const float4* qv, const float4* tp, const uint4* ctr, float2* qpos, float2* qvel,
                            float2* target, float2* tip, int32_t* step, uint32_t* episode, float2* qpos_lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = qv[i], b = tp[i];
    const uint4 c = ctr[i];
    if (qpos) qpos[i] = make_float2(a.x, a.y);
    if (qpos_lo) qpos_lo[i] = make_float2(__uint_as_float(c.z), __uint_as_float(c.w));
    if (qvel) qvel[i] = make_float2(a.z, a.w);
    if (target) target[i] = make_float2(b.x, b.y);
    if (tip) tip[i] = make_float2(b.z, b.w);
    if (step) step[i] = (int32_t)c.x;
    if (episode) episode[i] = c.y;
}
__global__ void k_set_state(int64_t n, float4* qv, float4* tp, uint4* ctr, const float2* qpos, const float2* qvel,
                            const float2* target, const float2* tip, const int32_t* step, const uint32_t* episode, const float2* qpos_lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = qv[i], b = tp[i];
    uint4 c = ctr[i];
    if (qpos) { const float2 v = qpos[i]; a.x = v.x; a.y = v.y; c.z = c.w = 0u; }       // new angles: low parts start at zero ...
    if (qpos_lo) { const float2 v = qpos_lo[i]; c.z = __float_as_uint(v.x); c.w = __float_as_uint(v.y); }   // ... unless given (exact resume)
    if (qvel) { const float2 v = qvel[i]; a.z = v.x; a.w = v.y; }
    if (target) { const float2 v = target[i]; b.x = v.x; b.y = v.y; }
    if (tip) { const float2 v = tip[i]; b.z = v.x; b.w = v.y; }
    else if (qpos) fk(a.x, a.y, b.z, b.w);
    if (step) c.x = (uint32_t)step[i];
    if (episode) c.y = episode[i];
    qv[i] = a; tp[i] = b; ctr[i] = c;
}

// Fused T-step rollout with Philox random actions: state lives in registers for all T steps (FP32-pipe bound).
__global__ void __launch_bounds__(ENV_BLOCK) k_rollout_random(int64_t n, float4* qv, float4* tp, uint4* ctr, int T, uint32_t step0,
                                                              float* __restrict__ obs_buf, float2* __restrict__ act_buf,
                                                              float* __restrict__ rew_buf, uint8_t* __restrict__ done_buf,
                                                              uint32_t k0, uint32_t k1, uint32_t offset) {
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * ENV_BLOCK + threadIdx.x;
    const int64_t row0 = i - lane;
    if (row0 >= n) return;
    const int nvalid = (int)min((int64_t)32, n - row0);
    const bool valid = i < n;
    const uint32_t gid = offset + (uint32_t)i;
    EnvState e;
    if (valid) e = load_state(qv, tp, ctr, i);
    else e = zero_state();
    for (int t = 0; t < T; ++t) {
        const uint4 r = philox4x32_10(gid, step0 + (uint32_t)t, 0u, STREAM_ACTION, k0, k1);
        const float a0 = uniform_f32(r.x, -1.f, 1.f), a1 = uniform_f32(r.y, -1.f, 1.f);
        bool d;
        const float rw = step_env(e, a0, a1, k0, k1, gid, d);
        const int64_t row = (int64_t)t * n + i;
        if (valid) {
            if (act_buf) act_buf[row] = make_float2(a0, a1);
            if (rew_buf) rew_buf[row] = rw;
            if (done_buf) done_buf[row] = d ? 1 : 0;
        }
        if (obs_buf) {
            float ob[OBS];
            observe(e, ob);
            warp_store_rows<OBS>(obs_buf, (int64_t)t * n + row0, nvalid, ob, strips[warp], lane);
        }
    }
    if (valid) store_state(qv, tp, ctr, i, e);
}

// Standalone policy forward (one thread per sample)
template <int NOUT>
__global__ void __launch_bounds__(ENV_BLOCK) k_policy_fwd_fp32(const float* __restrict__ params, const float* __restrict__ obs,
                                                               int64_t n, float4* __restrict__ pd) {
    __shared__ PolicySmem S;
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    policy_load_smem(S, params, NOUT);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = (int64_t)blockIdx.x * ENV_BLOCK; base < n; base += (int64_t)gridDim.x * ENV_BLOCK) {
        const int64_t i = base + threadIdx.x;
        const int64_t row0 = i - lane;
        if (row0 >= n) continue;
        const int nvalid = (int)min((int64_t)32, n - row0);
        float ob[OBS], o[4];
        warp_load_rows<OBS>(obs, row0, nvalid, ob, strips[warp], lane);
        policy_fwd_simt<NOUT>(S, ob, o);
        if (i < n) pd[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// Fused policy-in-the-loop rollout (teacher warm-up loop, mlp_train.py:120-139), fp32 CUDA-core policy.
template <int NOUT>
__global__ void __launch_bounds__(ENV_BLOCK) k_rollout_policy_fp32(int64_t n, float4* qv, float4* tp, uint4* ctr, const float* __restrict__ params,
                                                                   int T, float* __restrict__ obs_buf, float4* __restrict__ pd_buf,
                                                                   float* __restrict__ rew_buf, uint8_t* __restrict__ done_buf,
                                                                   uint32_t k0, uint32_t k1, uint32_t offset, uint64_t* __restrict__ done_mask,
                                                                   float* __restrict__ return_sum) {
    __shared__ PolicySmem S;
    __shared__ __align__(16) float strips[ENV_BLOCK / 32][32 * OBS];
    policy_load_smem(S, params, NOUT);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * ENV_BLOCK + threadIdx.x;
    const int64_t row0 = i - lane;
    if (row0 >= n) return;
    const int nvalid = (int)min((int64_t)32, n - row0);
    const bool valid = i < n;
    const uint32_t gid = offset + (uint32_t)i;
    EnvState e;
    if (valid) e = load_state(qv, tp, ctr, i);
    else e = zero_state();
    uint64_t dmask = 0ull;
    float rsum = 0.f;
    for (int t = 0; t < T; ++t) {
        float ob[OBS], pd[4];
        observe(e, ob);
        policy_fwd_simt<NOUT>(S, ob, pd);
        bool d;
        const float rw = step_env(e, pd[0], pd[1], k0, k1, gid, d);
        dmask |= (uint64_t)(d ? 1u : 0u) << (t & 63);
        rsum = __fadd_rn(rsum, rw);
        const int64_t row = (int64_t)t * n + i;
        if (valid) {
            if (pd_buf) pd_buf[row] = make_float4(pd[0], pd[1], pd[2], pd[3]);
            if (rew_buf) rew_buf[row] = rw;
            if (done_buf) done_buf[row] = d ? 1 : 0;
        }
        if (obs_buf) warp_store_rows<OBS>(obs_buf, (int64_t)t * n + row0, nvalid, ob, strips[warp], lane);
    }
    if (valid) { store_state(qv, tp, ctr, i, e); if (done_mask) done_mask[i] = dmask; if (return_sum) return_sum[i] = rsum; }
}

inline unsigned env_grid(int64_t n) { return (unsigned)((n + ENV_BLOCK - 1) / ENV_BLOCK); }

// implemented in policy_tc.cu
int policy_fwd_tc(const float* params, int nout, const float* obs, int64_t n, float* pd, cudaStream_t s);
int rollout_policy_tc(rb_env* env, const float* params, int nout, int T, float* obs_buf, float* pd_buf, float* rew_buf,
                      uint8_t* done_buf, cudaStream_t s);

}  // namespace rb

using namespace rb;

extern "C" {

int rb_env_create(rb_env** out, int64_t num_envs, uint64_t seed, int device, uint32_t global_env_offset) {
    RB_REQUIRE(out != nullptr, "out is NULL");
    RB_REQUIRE(num_envs > 0 && num_envs <= (int64_t)1 << 31, "num_envs out of range");
    RB_REQUIRE((uint64_t)global_env_offset + (uint64_t)num_envs <= (uint64_t)1 << 32, "global env id overflows 32 bits");
    DeviceGuard guard(device);             // the caller's current device is put back on return
    RB_CUDA(cudaGetLastError());
    rb_env* e = new rb_env();
    e->n = num_envs; e->seed = seed; e->device = device; e->offset = global_env_offset;
    cudaError_t err = cudaMalloc(&e->qv, sizeof(float4) * num_envs);
    if (err == cudaSuccess) err = cudaMalloc(&e->tp, sizeof(float4) * num_envs);
    if (err == cudaSuccess) err = cudaMalloc(&e->ctr, sizeof(uint4) * num_envs);
    if (err == cudaSuccess) err = cudaMemset(e->qv, 0, sizeof(float4) * num_envs);
    if (err == cudaSuccess) err = cudaMemset(e->tp, 0, sizeof(float4) * num_envs);
    if (err == cudaSuccess) err = cudaMemset(e->ctr, 0, sizeof(uint4) * num_envs);
    cudaDeviceProp prop;
    if (err == cudaSuccess) err = cudaGetDeviceProperties(&prop, device);
    if (err != cudaSuccess) { rb_env_destroy(e); return cuda_fail(err, "rb_env_create"); }
    e->sm_count = prop.multiProcessorCount;
    *out = e;
    return RB_OK;
}

int rb_env_destroy(rb_env* e) {
    if (!e) return RB_OK;
    env_serve_destroy(e);
    DeviceGuard guard(e->device);
    cudaFree(e->qv); cudaFree(e->tp); cudaFree(e->ctr);
    cudaFree(e->d_act); cudaFree(e->d_obs); cudaFree(e->d_rew); cudaFree(e->d_done); cudaFree(e->d_params);
    cudaFree(e->d_buf_obs); cudaFree(e->d_buf_pd); cudaFree(e->d_buf_rew); cudaFree(e->d_buf_done);
    cudaFree(e->prog_counters); cudaFree(e->d_done_mask); cudaFree(e->d_return_sum); cudaFree(e->d_params2); cudaFree(e->d_rew_pipe[0]); cudaFree(e->d_rew_pipe[1]);
    for (int i = 0; i < 2; ++i) if (e->pipe_kernel[i]) cudaEventDestroy(e->pipe_kernel[i]);
    for (int i = 0; i < 2; ++i) if (e->pipe_done[i]) cudaEventDestroy(e->pipe_done[i]);
    if (e->prog_flags_host) cudaFreeHost((void*)e->prog_flags_host);
    if (e->host_stream) cudaStreamDestroy(e->host_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int i = 0; i < 16; ++i) if (e->slab_done[i]) cudaEventDestroy(e->slab_done[i]);
    delete e;
    return RB_OK;
}

int64_t rb_env_num_envs(const rb_env* e) { return e ? e->n : 0; }

int rb_env_reset(rb_env* e, float* obs_dev, void* stream) {
    RB_REQUIRE(e != nullptr, "env is NULL");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_reset<<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, obs_dev, (uint32_t)e->seed,
                                                                    (uint32_t)(e->seed >> 32), e->offset);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_env_observe(rb_env* e, float* obs_dev, void* stream) {
    RB_REQUIRE(e != nullptr && obs_dev != nullptr, "NULL argument");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_observe<<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, obs_dev);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_env_step(rb_env* e, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream) {
    RB_REQUIRE(e != nullptr && act_dev != nullptr, "NULL argument");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_step<<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, (const float2*)act_dev, obs_dev, rew_dev,
                                                                   done_dev, (uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->offset);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

static int ensure_host_staging(rb_env* e) {
    if (e->d_act) return RB_OK;
    DeviceGuard guard(e->device);
    RB_CUDA(cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking));
    RB_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) RB_CUDA(cudaEventCreateWithFlags(&e->slab_done[i], cudaEventDisableTiming));
    RB_CUDA(cudaMalloc(&e->d_act, sizeof(float) * 2 * e->n));
    RB_CUDA(cudaMalloc(&e->d_obs, sizeof(float) * OBS * e->n));
    RB_CUDA(cudaMalloc(&e->d_rew, sizeof(float) * e->n));
    RB_CUDA(cudaMalloc(&e->d_done, e->n));
    RB_CUDA(cudaMalloc(&e->d_params, sizeof(float) * rb_policy_param_count(4)));
    RB_CUDA(cudaMalloc(&e->prog_counters, 16 * sizeof(uint32_t)));
    void* h = nullptr;
    RB_CUDA(cudaHostAlloc(&h, 16 * sizeof(uint32_t), cudaHostAllocMapped));
    memset(h, 0, 16 * sizeof(uint32_t));
    void* dv = nullptr;
    RB_CUDA(cudaHostGetDevicePointer(&dv, h, 0));
    e->prog_flags_host = (volatile uint32_t*)h; e->prog_flags_dev = (uint32_t*)dv;
    return RB_OK;
}

int rb_env_reset_host(rb_env* e, float* obs_host) {
    RB_REQUIRE(e != nullptr && obs_host != nullptr, "NULL argument");
    if (env_serve_eligible(e)) return env_serve_reset(e, obs_host, nullptr);      // small host-surface envs: resident server, no launch (serve.cu)
    int rc = ensure_host_staging(e);
    if (rc) return rc;
    rc = rb_env_reset(e, e->d_obs, e->host_stream);
    if (rc) return rc;
    RB_CUDA(cudaMemcpyAsync(obs_host, e->d_obs, sizeof(float) * OBS * e->n, cudaMemcpyDeviceToHost, e->host_stream));
    RB_CUDA(cudaStreamSynchronize(e->host_stream));
    return RB_OK;
}

int rb_env_step_host(rb_env* e, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host) {
    RB_REQUIRE(e != nullptr && act_host != nullptr && obs_host != nullptr, "NULL argument");
    if (env_serve_eligible(e)) return env_serve_step(e, act_host, obs_host, rew_host, done_host, nullptr);
    int rc = ensure_host_staging(e);
    if (rc) return rc;
    cudaStream_t s = e->host_stream;
    RB_CUDA(cudaMemcpyAsync(e->d_act, act_host, sizeof(float) * 2 * e->n, cudaMemcpyHostToDevice, s));
    rc = rb_env_step(e, e->d_act, e->d_obs, e->d_rew, e->d_done, s);
    if (rc) return rc;
    RB_CUDA(cudaMemcpyAsync(obs_host, e->d_obs, sizeof(float) * OBS * e->n, cudaMemcpyDeviceToHost, s));
    if (rew_host) RB_CUDA(cudaMemcpyAsync(rew_host, e->d_rew, sizeof(float) * e->n, cudaMemcpyDeviceToHost, s));
    if (done_host) RB_CUDA(cudaMemcpyAsync(done_host, e->d_done, e->n, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

int rb_env_get_state(rb_env* e, float* qpos, float* qvel, float* target, float* tip, int32_t* step, uint32_t* episode, float* qpos_lo,
                     void* stream) {
    RB_REQUIRE(e != nullptr, "env is NULL");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_get_state<<<(unsigned)((e->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, (float2*)qpos, (float2*)qvel,
                                                                                 (float2*)target, (float2*)tip, step, episode, (float2*)qpos_lo);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_env_set_state(rb_env* e, const float* qpos, const float* qvel, const float* target, const float* tip, const int32_t* step,
                     const uint32_t* episode, const float* qpos_lo, void* stream) {
    RB_REQUIRE(e != nullptr, "env is NULL");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_set_state<<<(unsigned)((e->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, (const float2*)qpos,
                                                                                 (const float2*)qvel, (const float2*)target,
                                                                                 (const float2*)tip, step, episode, (const float2*)qpos_lo);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_env_rollout_random(rb_env* e, int T, uint32_t step0, float* obs_buf, float* act_buf, float* rew_buf, uint8_t* done_buf, void* stream) {
    RB_REQUIRE(e != nullptr && T >= 0, "bad argument");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    if (T == 0) return RB_OK;
    k_rollout_random<<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, T, step0, obs_buf, (float2*)act_buf,
                                                                             rew_buf, done_buf, (uint32_t)e->seed,
                                                                             (uint32_t)(e->seed >> 32), e->offset);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int64_t rb_policy_param_count(int nout) { return (nout == 2 || nout == 4) ? policy_offsets(nout).total : -1; }

int rb_policy_fwd(const float* params, int nout, const float* obs, int64_t n, float* pd, int mode, void* stream) {
    RB_REQUIRE(params && obs && pd, "NULL argument");
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    RB_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return RB_OK;
    if (mode == RB_MODE_TC) return policy_fwd_tc(params, nout, obs, n, pd, (cudaStream_t)stream);
    RB_REQUIRE(mode == RB_MODE_FP32, "unknown mode");
    const unsigned grid = (unsigned)min((int64_t)148 * 8, (n + ENV_BLOCK - 1) / ENV_BLOCK);
    if (nout == 2) k_policy_fwd_fp32<2><<<grid, ENV_BLOCK, 0, (cudaStream_t)stream>>>(params, obs, n, (float4*)pd);
    else k_policy_fwd_fp32<4><<<grid, ENV_BLOCK, 0, (cudaStream_t)stream>>>(params, obs, n, (float4*)pd);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_policy_fwd_host(const float* params_host, int nout, const float* obs_host, int64_t n, float* pd_host, int mode, int device) {
    RB_REQUIRE(params_host && obs_host && pd_host, "NULL argument");
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    DeviceGuard guard(device);
    float *dp = nullptr, *dob = nullptr, *dpd = nullptr;
    const int64_t P = rb_policy_param_count(nout);
    RB_CUDA(cudaMalloc(&dp, sizeof(float) * P));
    RB_CUDA(cudaMalloc(&dob, sizeof(float) * OBS * (n > 0 ? n : 1)));
    RB_CUDA(cudaMalloc(&dpd, sizeof(float) * 4 * (n > 0 ? n : 1)));
    int rc = RB_OK;
    cudaError_t err = cudaMemcpy(dp, params_host, sizeof(float) * P, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(dob, obs_host, sizeof(float) * OBS * n, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) rc = rb_policy_fwd(dp, nout, dob, n, dpd, mode, nullptr);
    if (err == cudaSuccess && rc == RB_OK) err = cudaMemcpy(pd_host, dpd, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost);
    cudaFree(dp); cudaFree(dob); cudaFree(dpd);
    if (err != cudaSuccess) return cuda_fail(err, "rb_policy_fwd_host");
    return rc;
}

int rb_env_rollout_policy(rb_env* e, const float* params, int nout, int T, float* obs_buf, float* pd_buf, float* rew_buf,
                          uint8_t* done_buf, int mode, void* stream) {
    RB_REQUIRE(e != nullptr && params != nullptr && T >= 0, "bad argument");
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    if (T == 0) return RB_OK;
    if (mode == RB_MODE_TC) return rollout_policy_tc(e, params, nout, T, obs_buf, pd_buf, rew_buf, done_buf, (cudaStream_t)stream);
    RB_REQUIRE(mode == RB_MODE_FP32, "unknown mode");
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    if (nout == 2)
        k_rollout_policy_fp32<2><<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, params, T, obs_buf,
                                                                                         (float4*)pd_buf, rew_buf, done_buf, k0, k1, e->offset, e->done_mask_out, e->return_sum_out);
    else
        k_rollout_policy_fp32<4><<<env_grid(e->n), ENV_BLOCK, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, params, T, obs_buf,
                                                                                         (float4*)pd_buf, rew_buf, done_buf, k0, k1, e->offset, e->done_mask_out, e->return_sum_out);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

// device-visible alias of a HOST pointer when the memory is page-locked and mapped (cudaHostAlloc / cudaHostRegister under UVA), else NULL
static void* mapped_alias(const void* host) {
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}
// Transport constants of rb_env_rollout_policy_host (each measured against its alternatives on B200, see the comment inside the call and
// profiles/r01_e2e_transport_sweep.jsonl; they were environment knobs while being tuned):
constexpr int HOST_SLABS = 6;           // fp32-mode slab launches when bulk fields (obs / pdflat / pageable reward) go to the host
constexpr int HOST_SLAB_FIRST = 2;      // ... with a short first slab so that the copy engine starts early
constexpr int HOST_SLABS_SMALL = 1;     // ... and when only small fields are copied
constexpr int HOST_PROGRESS_SLABS = 5;  // tensor-core mode: time slabs reported by the ONE launch through in-kernel progress flags

static int rollout_policy_host_impl(rb_env* e, const float* params_host, int nout, int T, float* obs_host, float* pd_host, float* rew_host,
                                    uint8_t* done_host, int mode);

// host staging + the device-resident rollout buffer for T steps
static int ensure_rollout_buffer(rb_env* e, int T) {
    int rc = ensure_host_staging(e);
    if (rc) return rc;
    if (e->buf_T < T) {
        DeviceGuard guard(e->device);
        cudaFree(e->d_buf_obs); cudaFree(e->d_buf_pd); cudaFree(e->d_buf_rew); cudaFree(e->d_buf_done);
        e->d_buf_obs = e->d_buf_pd = e->d_buf_rew = nullptr; e->d_buf_done = nullptr; e->buf_T = 0;
        const int64_t rows = (int64_t)T * e->n;
        RB_CUDA(cudaMalloc(&e->d_buf_obs, sizeof(float) * OBS * rows));
        RB_CUDA(cudaMalloc(&e->d_buf_pd, sizeof(float) * 4 * rows));
        RB_CUDA(cudaMalloc(&e->d_buf_rew, sizeof(float) * rows));
        RB_CUDA(cudaMalloc(&e->d_buf_done, rows));
        e->buf_T = T;
    }
    e->buf_last_T = T;
    return RB_OK;
}

/* Split-phase form of the host rollout for callers that keep the GPU busy: _begin queues {H2D of the parameters, the rollout launch} and returns;
 * _wait blocks until the OLDEST outstanding call has finished (its outputs are then complete in host memory).  Up to two calls may be in flight,
 * so the host-side work around call i (parameter copy, launch, completion wait) overlaps the kernel of call i + 1.  Outputs are the kernel-stored
 * ones only -- reward [T,N], done_mask [N], return_sum [N], each NULL or a page-locked (mapped) buffer that stays untouched until its _wait. */
int rb_env_rollout_policy_host_begin(rb_env* e, const float* params_host, int nout, int T, float* rew_host, uint64_t* done_mask_host,
                                     float* return_sum_host, int mode) {
    RB_REQUIRE(e != nullptr && params_host != nullptr && T > 0, "bad argument");
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    RB_REQUIRE(!done_mask_host || T <= 64, "done_mask holds one bit per step: T <= 64");
    RB_REQUIRE(e->pipe_issued - e->pipe_waited < 2, "two rollouts are already in flight: call rb_env_rollout_policy_host_wait");
    float* zr = (float*)mapped_alias(rew_host);
    uint64_t* zm = (uint64_t*)mapped_alias(done_mask_host);
    float* zs = (float*)mapped_alias(return_sum_host);
    RB_REQUIRE((!rew_host || zr) && (!done_mask_host || zm) && (!return_sum_host || zs), "the split-phase rollout needs page-locked (mapped) output buffers");
    int rc = ensure_rollout_buffer(e, T);
    if (rc) return rc;
    DeviceGuard guard(e->device);
    const int slot = (int)(e->pipe_issued & 1);
    if (!e->d_params2) RB_CUDA(cudaMalloc(&e->d_params2, sizeof(float) * rb_policy_param_count(4)));
    if (!e->pipe_done[slot]) RB_CUDA(cudaEventCreateWithFlags(&e->pipe_done[slot], cudaEventDisableTiming));
    float* dp = slot ? e->d_params2 : e->d_params;             // the call still in flight reads the other copy
    cudaStream_t s = e->host_stream, sc = e->copy_stream;
    RB_CUDA(cudaMemcpyAsync(dp, params_host, sizeof(float) * rb_policy_param_count(nout), cudaMemcpyHostToDevice, s));
    // reward [T,N] (the bulk of the result): stored by the kernel straight into the host buffer (rb_env_set_host_transport bit 0, default), or
    // -- bit 0 clear -- written to a per-slot device buffer and brought over by the copy engine WHILE THE NEXT CALL'S KERNEL RUNS: the posted
    // PCIe stores slow the kernel by ~3 % (0.305 vs 0.295 ms at config 3), a copy of the previous chunk beside it does not.
    const bool rew_copy = rew_host && !(e->host_zerocopy & 1);
    float* rew_dst = zr ? zr : e->d_buf_rew;
    if (rew_copy) {
        const int64_t rows = (int64_t)T * e->n;
        if (e->rew_pipe_rows < rows) {
            RB_REQUIRE(e->pipe_issued == e->pipe_waited, "reward staging can only grow while no split-phase rollout is in flight");
            for (int i = 0; i < 2; ++i) { cudaFree(e->d_rew_pipe[i]); e->d_rew_pipe[i] = nullptr; RB_CUDA(cudaMalloc(&e->d_rew_pipe[i], sizeof(float) * rows)); }
            e->rew_pipe_rows = rows;
        }
        if (!e->pipe_kernel[slot]) RB_CUDA(cudaEventCreateWithFlags(&e->pipe_kernel[slot], cudaEventDisableTiming));
        rew_dst = e->d_rew_pipe[slot];
    }
    e->done_mask_out = zm; e->return_sum_out = zs;
    rc = rb_env_rollout_policy(e, dp, nout, T, e->d_buf_obs, e->d_buf_pd, rew_dst, nullptr, mode, s);
    e->done_mask_out = nullptr; e->return_sum_out = nullptr;
    if (rc) return rc;
    if (rew_copy) {
        RB_CUDA(cudaEventRecord(e->pipe_kernel[slot], s));
        RB_CUDA(cudaStreamWaitEvent(sc, e->pipe_kernel[slot], 0));
        RB_CUDA(cudaMemcpyAsync(rew_host, rew_dst, sizeof(float) * (size_t)T * e->n, cudaMemcpyDeviceToHost, sc));
        RB_CUDA(cudaEventRecord(e->pipe_done[slot], sc));      // the call is complete when its reward has landed (the mask / return words were stored by the kernel)
    } else {
        RB_CUDA(cudaEventRecord(e->pipe_done[slot], s));
    }
    e->pipe_issued += 1;
    return RB_OK;
}

int rb_env_rollout_policy_host_wait(rb_env* e) {
    RB_REQUIRE(e != nullptr, "env is NULL");
    RB_REQUIRE(e->pipe_waited < e->pipe_issued, "no split-phase rollout is in flight");
    RB_CUDA(cudaEventSynchronize(e->pipe_done[e->pipe_waited & 1]));
    e->pipe_waited += 1;
    return RB_OK;
}

int rb_env_rollout_policy_host(rb_env* e, const float* params_host, int nout, int T, float* obs_host, float* pd_host, float* rew_host,
                               uint8_t* done_host, int mode) {
    return rb_env_rollout_policy_host_ex(e, params_host, nout, T, obs_host, pd_host, rew_host, done_host, nullptr, nullptr, mode);
}

int rb_env_rollout_policy_host_ex(rb_env* e, const float* params_host, int nout, int T, float* obs_host, float* pd_host, float* rew_host,
                                  uint8_t* done_host, uint64_t* done_mask_host, float* return_sum_host, int mode) {
    RB_REQUIRE(e != nullptr && params_host != nullptr && T > 0, "bad argument");
    if (!done_mask_host && !return_sum_host) return rollout_policy_host_impl(e, params_host, nout, T, obs_host, pd_host, rew_host, done_host, mode);
    RB_REQUIRE(!done_mask_host || T <= 64, "done_mask holds one bit per step: T <= 64");
    // page-locked + mapped buffers: the kernel stores the per-env words straight into them; otherwise device staging + one copy behind the call
    uint64_t* zm = (uint64_t*)mapped_alias(done_mask_host);
    float* zr = (float*)mapped_alias(return_sum_host);
    DeviceGuard guard(e->device);
    if (done_mask_host && !zm && !e->d_done_mask) RB_CUDA(cudaMalloc(&e->d_done_mask, sizeof(uint64_t) * e->n));
    if (return_sum_host && !zr && !e->d_return_sum) RB_CUDA(cudaMalloc(&e->d_return_sum, sizeof(float) * e->n));
    e->done_mask_out = done_mask_host ? (zm ? zm : e->d_done_mask) : nullptr;
    e->return_sum_out = return_sum_host ? (zr ? zr : e->d_return_sum) : nullptr;
    const int rc = rollout_policy_host_impl(e, params_host, nout, T, obs_host, pd_host, rew_host, done_host, mode);
    e->done_mask_out = nullptr; e->return_sum_out = nullptr;
    if (rc) return rc;
    if (done_mask_host && !zm) RB_CUDA(cudaMemcpy(done_mask_host, e->d_done_mask, sizeof(uint64_t) * e->n, cudaMemcpyDeviceToHost));
    if (return_sum_host && !zr) RB_CUDA(cudaMemcpy(return_sum_host, e->d_return_sum, sizeof(float) * e->n, cudaMemcpyDeviceToHost));
    return RB_OK;
}

static int rollout_policy_host_impl(rb_env* e, const float* params_host, int nout, int T, float* obs_host, float* pd_host, float* rew_host,
                                    uint8_t* done_host, int mode) {
    RB_REQUIRE(e != nullptr && params_host != nullptr && T > 0, "bad argument");
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    RB_REQUIRE(e->pipe_issued == e->pipe_waited, "a split-phase rollout (rb_env_rollout_policy_host_begin) is still in flight: wait for it first");
    int rc = ensure_rollout_buffer(e, T);
    if (rc) return rc;
    cudaStream_t s = e->host_stream, sc = e->copy_stream;
    RB_CUDA(cudaMemcpyAsync(e->d_params, params_host, sizeof(float) * rb_policy_param_count(nout), cudaMemcpyHostToDevice, s));
    // The kernel ALWAYS fills the device-resident rollout buffer (obs, pdflat: rb_env_rollout_buffer() hands it to the distillation loop);
    // a NULL host pointer only means "do not bring that field to the host".  How the fields reach the host:
    //  * reward, when the host buffer is page-locked and mapped: the kernel stores it straight into host memory (one posted 128-byte PCIe
    //    write per warp-step, issued while the rollout runs -- 13 MB at config 3, hidden under the kernel).  RB_HOST_ZEROCOPY (bit 0: reward,
    //    bit 1: done; default 1) selects this; `done` is 32 bytes per warp-step, which makes poor PCIe packets, so by default it takes
    //    the copy path;
    //  * everything else goes through the device buffer and the copy engine in TIME SLABS (same trajectories: the state round-trips
    //    through HBM exactly): the device->host copy of slab i (copy stream) overlaps the kernel of slab i+1 (compute stream).  A slab
    //    boundary costs ~10 us of kernel time and its copy competes with the kernel's own PCIe writes, so when only small fields are
    //    left to copy they go in RB_HOST_SLABS_SMALL (1) slab(s) after the kernel; obs / pdflat / pageable reward (PCIe-bound bulk) use
    //    RB_HOST_SLABS (6) slabs with a short first one (RB_HOST_SLAB_FIRST, 2 steps) so that the copy engine starts early.
    //  Tensor-core mode does not cut the launch at all: see the in-kernel progress path below.
    //  Measured on B200, 65 536 envs x 50 steps, result = reward + done, ms per call (scripts/e2e_sweep.py; kernel alone 0.307, the 16.4 MB
    //  result alone 0.30 at the 55 GB/s this box copies at): 0.362 progress path with the defaults (reward kernel-written, done copied per
    //  progress slab); 0.435 progress path with everything on the copy engine; 0.403 reward kernel-written + done copied after the kernel;
    //  0.419 the same with 2 kernel slabs; 0.428 both fields kernel-written; 0.448 everything copied in 5 equal kernel slabs.
    const int zc = e->host_zerocopy;
    constexpr int nslab_bulk = HOST_SLABS, first_steps = HOST_SLAB_FIRST, nslab_small = HOST_SLABS_SMALL;
    float* rew_zc = (zc & 1) ? (float*)mapped_alias(rew_host) : nullptr;
    uint8_t* done_zc = (zc & 2) ? (uint8_t*)mapped_alias(done_host) : nullptr;
    const bool copy_rew = rew_host && !rew_zc, copy_done = done_host && !done_zc;
    const bool bulk = obs_host || pd_host || copy_rew, any_copy = bulk || copy_done;
    if (mode == RB_MODE_TC && any_copy) {
        // Tensor-core rollout: ONE launch for all T steps; the kernel posts per-slab completion flags into mapped host memory (common.cuh:
        // rb_env::prog_*) and this thread, polling them, queues the copy of slab i on the copy stream while the launch computes slab i+1.
        // No slab boundaries in the kernel (each costs ~10 us), only the last slab's copy is left after it.
        constexpr int nprog = HOST_PROGRESS_SLABS;
        const int slab_len = (T + nprog - 1) / nprog, nsl = (T + slab_len - 1) / slab_len;
        e->prog_epoch += 1u;
        RB_CUDA(cudaMemsetAsync(e->prog_counters, 0, 16 * sizeof(uint32_t), s));
        e->prog_slab_len = slab_len;
        rc = rb_env_rollout_policy(e, e->d_params, nout, T, e->d_buf_obs, e->d_buf_pd, rew_zc ? rew_zc : e->d_buf_rew, done_zc ? done_zc : e->d_buf_done, mode, s);
        e->prog_slab_len = 0;
        if (rc) return rc;
        const auto t_start = std::chrono::steady_clock::now();
        for (int i = 0; i < nsl; ++i) {
            uint64_t spins = 0;
            while (e->prog_flags_host[i] != e->prog_epoch) {
                if ((++spins & 0x3FFu) == 0) {
                    const cudaError_t q = cudaStreamQuery(s);
                    if (q == cudaSuccess) break;                                   // the launch is over: everything is written
                    if (q != cudaErrorNotReady) return cuda_fail(q, "rb_env_rollout_policy_host");
                    if (std::chrono::steady_clock::now() - t_start > std::chrono::seconds(60)) { set_error("rb_env_rollout_policy_host: timed out"); return RB_ERR_CUDA; }
                }
            }
            const int t0 = i * slab_len, tn = (t0 + slab_len <= T ? slab_len : T - t0);
            const int64_t r0 = (int64_t)t0 * e->n, rows = (int64_t)tn * e->n;
            if (copy_done) RB_CUDA(cudaMemcpyAsync(done_host + r0, e->d_buf_done + r0, rows, cudaMemcpyDeviceToHost, sc));
            if (copy_rew) RB_CUDA(cudaMemcpyAsync(rew_host + r0, e->d_buf_rew + r0, sizeof(float) * rows, cudaMemcpyDeviceToHost, sc));
            if (obs_host) RB_CUDA(cudaMemcpyAsync(obs_host + OBS * r0, e->d_buf_obs + OBS * r0, sizeof(float) * OBS * rows, cudaMemcpyDeviceToHost, sc));
            if (pd_host) RB_CUDA(cudaMemcpyAsync(pd_host + 4 * r0, e->d_buf_pd + 4 * r0, sizeof(float) * 4 * rows, cudaMemcpyDeviceToHost, sc));
        }
        RB_CUDA(cudaStreamSynchronize(sc));
        RB_CUDA(cudaStreamSynchronize(s));
        // Measured (config 3, B200): flags at 75 / 138 / 195 / 257 / 316 us, the last copy done at ~340, call returns at ~350 (+ ~12 us before the
        // launch).  Tried on top and dropped (no gain, PCIe is shared by the kernel's reward stores and the copy engine): `done` of the last slab
        // kernel-written too (last slab 20 us slower, 0.367 ms per call), kernel sync before copy sync, self-resetting counters (0.3626).
        return RB_OK;
    }
    int ends[16];                                                                    // cumulative slab ends
    int nslab = (!any_copy || e->done_mask_out || e->return_sum_out) ? 1 : (bulk ? nslab_bulk : nslab_small);      // the per-env done mask counts steps from the launch's first step: one launch
    if (nslab > T) nslab = T;
    if (bulk) {
        const int first = (nslab > 1 && first_steps < T / nslab) ? first_steps : 0;  // 0: equal slabs
        for (int i = 0; i < nslab; ++i) ends[i] = first ? first + (int)(((int64_t)(T - first) * i) / (nslab - 1)) : (int)(((int64_t)T * (i + 1)) / nslab);
    } else {
        for (int i = 0; i < nslab; ++i) ends[i] = T - (T >> (2 * (i + 1)));          // 3/4, 15/16, ... of T
    }
    ends[nslab - 1] = T;
    int t0 = 0;
    for (int i = 0; i < nslab; ++i) {
        const int tn = ends[i] - t0;
        if (tn <= 0) continue;
        const int64_t r0 = (int64_t)t0 * e->n, rows = (int64_t)tn * e->n;
        rc = rb_env_rollout_policy(e, e->d_params, nout, tn, e->d_buf_obs + OBS * r0, e->d_buf_pd + 4 * r0, rew_zc ? rew_zc + r0 : e->d_buf_rew + r0,
                                   done_zc ? done_zc + r0 : e->d_buf_done + r0, mode, s);
        if (rc) return rc;
        if (any_copy) {
            RB_CUDA(cudaEventRecord(e->slab_done[i], s));
            RB_CUDA(cudaStreamWaitEvent(sc, e->slab_done[i], 0));
            if (copy_done) RB_CUDA(cudaMemcpyAsync(done_host + r0, e->d_buf_done + r0, rows, cudaMemcpyDeviceToHost, sc));
            if (copy_rew) RB_CUDA(cudaMemcpyAsync(rew_host + r0, e->d_buf_rew + r0, sizeof(float) * rows, cudaMemcpyDeviceToHost, sc));
            if (obs_host) RB_CUDA(cudaMemcpyAsync(obs_host + OBS * r0, e->d_buf_obs + OBS * r0, sizeof(float) * OBS * rows, cudaMemcpyDeviceToHost, sc));
            if (pd_host) RB_CUDA(cudaMemcpyAsync(pd_host + 4 * r0, e->d_buf_pd + 4 * r0, sizeof(float) * 4 * rows, cudaMemcpyDeviceToHost, sc));
        }
        t0 += tn;
    }
    if (any_copy) RB_CUDA(cudaStreamSynchronize(sc));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

int rb_env_set_host_transport(rb_env* e, int kernel_stores) {
    RB_REQUIRE(e != nullptr && kernel_stores >= 0 && kernel_stores <= 3, "bad argument");
    e->host_zerocopy = kernel_stores;
    return RB_OK;
}

int rb_env_rollout_buffer(rb_env* e, float** obs_dev, float** pd_dev, float** rew_dev, uint8_t** done_dev, int* T) {
    RB_REQUIRE(e != nullptr, "env is NULL");
    RB_REQUIRE(e->buf_last_T > 0, "no host rollout has run on this env yet");
    if (obs_dev) *obs_dev = e->d_buf_obs;
    if (pd_dev) *pd_dev = e->d_buf_pd;
    if (rew_dev) *rew_dev = e->d_buf_rew;
    if (done_dev) *done_dev = e->d_buf_done;
    if (T) *T = (int)e->buf_last_T;
    return RB_OK;
}

}  // extern "C"
