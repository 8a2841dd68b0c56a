// Lock-step DAgger iteration pieces: observe (ob, teacher label, student input) and act (step with the student mean).
// Replaces the per-env-step body of /root/reference src/distilation/mlp_train.py:143-204 (teacher label :165-167, record
// :188-193 with dataset.py:118-143 `prev` / `prew` semantics, env.step(s_ac) :196) for N envs at once.
#include "common.cuh"
#include "physics.cuh"

struct rb_dagger {
    rb_env* env = nullptr;
    int kind = 0;
    float keep_prob = 1.f;
    float4* prev_t = nullptr;      // teacher pdflat of the previous record of the current episode
    float* prev_rec_rew = nullptr; // 'rew' field of the previous record
    float* last_reward = nullptr;  // reward returned by the last env.step (not cleared at reset, mlp_train.py:114,196)
};

namespace rb {

__global__ void k_dagger_input(int64_t n, const uint2* __restrict__ ctr, const float* __restrict__ obs, const float4* __restrict__ prev_t,
                               const float* __restrict__ prev_rec_rew, float keep_prob, uint32_t k0, uint32_t k1, uint32_t offset,
                               uint32_t iteration, float4* __restrict__ x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ob[12];
#pragma unroll
    for (int k = 0; k < 11; ++k) ob[k] = __ldg(obs + i * 11 + k);
    ob[11] = 0.f;
    if (keep_prob < 1.f) {
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            const uint4 r = philox4x32_10(offset + (uint32_t)i, iteration, (uint32_t)blk, STREAM_DROPOUT, k0, k1);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float u = (float)(rr[c] >> 8) * 5.9604644775390625e-08f;
                ob[4 * blk + c] = __fdiv_rn(ob[4 * blk + c], keep_prob) * floorf(keep_prob + u);
            }
        }
    }
    const bool first = ctr[i].x == 0u;   // first record of an episode: prev / prew are zeros (dataset.py:151-164)
    const float4 pp = first ? make_float4(0.f, 0.f, 0.f, 0.f) : prev_t[i];
    const float pr = first ? 0.f : prev_rec_rew[i];
    x[i * 4 + 0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
    x[i * 4 + 1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
    x[i * 4 + 2] = make_float4(ob[8], ob[9], ob[10], pp.x);
    x[i * 4 + 3] = make_float4(pp.y, pp.z, pp.w, pr);
}

__global__ void __launch_bounds__(128) k_dagger_act(int64_t n, float4* qv, float4* tp, uint2* ctr, const float4* __restrict__ s_pd,
                                                    const float4* __restrict__ t_pd, float4* prev_t, float* prev_rec_rew, float* last_reward,
                                                    float* __restrict__ rew, uint8_t* __restrict__ done, uint32_t k0, uint32_t k1,
                                                    uint32_t offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState e = load_state(qv, tp, ctr, i);
    const float4 sp = __ldg(s_pd + i);
    bool d;
    const float r = step_env(e, sp.x, sp.y, k0, k1, offset + (uint32_t)i, d);
    store_state(qv, tp, ctr, i, e);
    if (t_pd) prev_t[i] = __ldg(t_pd + i);
    prev_rec_rew[i] = last_reward[i];
    last_reward[i] = r;
    if (rew) rew[i] = r;
    if (done) done[i] = d ? 1 : 0;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_dagger_create(rb_dagger** out, rb_env* env, int student_kind, float keep_prob) {
    RB_REQUIRE(out && env, "NULL argument");
    RB_REQUIRE(student_kind == RB_STUDENT_POLICY64 || student_kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(keep_prob > 0.f, "keep_prob must be > 0");
    RB_CUDA(cudaSetDevice(env->device));
    rb_dagger* d = new rb_dagger();
    d->env = env; d->kind = student_kind; d->keep_prob = keep_prob;
    cudaError_t err = cudaMalloc(&d->prev_t, sizeof(float4) * env->n);
    if (err == cudaSuccess) err = cudaMalloc(&d->prev_rec_rew, sizeof(float) * env->n);
    if (err == cudaSuccess) err = cudaMalloc(&d->last_reward, sizeof(float) * env->n);
    if (err == cudaSuccess) err = cudaMemset(d->prev_t, 0, sizeof(float4) * env->n);
    if (err == cudaSuccess) err = cudaMemset(d->prev_rec_rew, 0, sizeof(float) * env->n);
    if (err == cudaSuccess) err = cudaMemset(d->last_reward, 0, sizeof(float) * env->n);
    if (err != cudaSuccess) { rb_dagger_destroy(d); return cuda_fail(err, "rb_dagger_create"); }
    *out = d;
    return RB_OK;
}

int rb_dagger_destroy(rb_dagger* d) {
    if (!d) return RB_OK;
    cudaFree(d->prev_t); cudaFree(d->prev_rec_rew); cudaFree(d->last_reward);
    delete d;
    return RB_OK;
}

int rb_dagger_observe(rb_dagger* d, const float* teacher_params, uint32_t iteration, float* obs, float* t_pd, float* x, int mode, void* stream) {
    RB_REQUIRE(d && teacher_params && obs && t_pd && x, "NULL argument");
    rb_env* e = d->env;
    int rc = rb_env_observe(e, obs, stream);
    if (rc) return rc;
    rc = rb_policy_fwd(teacher_params, 2, obs, e->n, t_pd, mode, stream);
    if (rc) return rc;
    if (d->kind == RB_STUDENT_POLICY64) {
        if (x != obs) RB_CUDA(cudaMemcpyAsync(x, obs, sizeof(float) * OBS * e->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    } else {
        k_dagger_input<<<(unsigned)((e->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->n, e->ctr, obs, d->prev_t, d->prev_rec_rew, d->keep_prob,
                                                                                        (uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->offset,
                                                                                        iteration, (float4*)x);
        RB_CUDA(cudaGetLastError());
    }
    return RB_OK;
}

int rb_dagger_act(rb_dagger* d, const float* s_pd, const float* t_pd, float* rew, uint8_t* done, void* stream) {
    RB_REQUIRE(d && s_pd, "NULL argument");
    rb_env* e = d->env;
    k_dagger_act<<<(unsigned)((e->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, (const float4*)s_pd, (const float4*)t_pd,
                                                                                  d->prev_t, d->prev_rec_rew, d->last_reward, rew, done,
                                                                                  (uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->offset);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // extern "C"
