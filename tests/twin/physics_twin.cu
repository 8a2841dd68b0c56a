// TEST INFRASTRUCTURE (not product code): the device physics source (reacherdistilation_b200/csrc/physics.cuh) compiled for the HOST.
// Every rounding in that header is explicit (fmaf / IEEE single operations, built with -ffp-contract=off), and the dynamics use no
// approximate instruction, so this library reproduces the device trajectories BIT FOR BIT (the reward differs by the MUFU sqrt's last
// ulp).  Two uses:  tests/test_physics_twin_cpu.py states the fp32 tolerance against the float64 oracle on the CPU, at BASELINE
// config 2's full size;  tests/test_env_gpu.py checks the kernels against it bit-exactly on the GPU box.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../reacherdistilation_b200/csrc/physics.cuh"

using namespace rb;

struct Twin {
    int64_t n;
    uint64_t seed;
    uint32_t offset;
    std::vector<EnvState> e;
};

extern "C" {

void* twin_create(int64_t n, uint64_t seed, uint32_t offset) {
    Twin* t = new Twin();
    t->n = n; t->seed = seed; t->offset = offset;
    t->e.assign((size_t)n, zero_state());
    return t;
}
void twin_destroy(void* h) { delete (Twin*)h; }

void twin_reset(void* h, float* obs) {
    Twin* t = (Twin*)h;
    const uint32_t k0 = (uint32_t)t->seed, k1 = (uint32_t)(t->seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < t->n; ++i) {
        EnvState& e = t->e[(size_t)i];
        e.episode = 0u;
        reset_env(e, k0, k1, t->offset + (uint32_t)i);
        if (obs) observe(e, obs + i * 11);
    }
}

void twin_step(void* h, const float* act, float* obs, float* rew, uint8_t* done) {
    Twin* t = (Twin*)h;
    const uint32_t k0 = (uint32_t)t->seed, k1 = (uint32_t)(t->seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < t->n; ++i) {
        EnvState& e = t->e[(size_t)i];
        bool d;
        const float r = step_env(e, act[2 * i], act[2 * i + 1], k0, k1, t->offset + (uint32_t)i, d);
        if (rew) rew[i] = r;
        if (done) done[i] = d ? 1 : 0;
        if (obs) observe(e, obs + i * 11);
    }
}

// T steps with the Philox action stream of rb_env_rollout_random (keyed (seed, global env, step0 + t)); obs_buf [T,N,11] etc. may be NULL
void twin_rollout_random(void* h, int T, uint32_t step0, float* obs_buf, float* act_buf, float* rew_buf, uint8_t* done_buf) {
    Twin* t = (Twin*)h;
    const uint32_t k0 = (uint32_t)t->seed, k1 = (uint32_t)(t->seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < t->n; ++i) {
        EnvState& e = t->e[(size_t)i];
        const uint32_t gid = t->offset + (uint32_t)i;
        for (int s = 0; s < T; ++s) {
            const uint4 r = philox4x32_10(gid, step0 + (uint32_t)s, 0u, STREAM_ACTION, k0, k1);
            const float a0 = uniform_f32(r.x, -1.f, 1.f), a1 = uniform_f32(r.y, -1.f, 1.f);
            bool d;
            const float rw = step_env(e, a0, a1, k0, k1, gid, d);
            const int64_t row = (int64_t)s * t->n + i;
            if (act_buf) { act_buf[2 * row] = a0; act_buf[2 * row + 1] = a1; }
            if (rew_buf) rew_buf[row] = rw;
            if (done_buf) done_buf[row] = d ? 1 : 0;
            if (obs_buf) observe(e, obs_buf + row * 11);
        }
    }
}

void twin_get_state(void* h, float* qpos, float* qvel, float* target, float* tip, int32_t* step, uint32_t* episode, float* qpos_lo) {
    Twin* t = (Twin*)h;
    for (int64_t i = 0; i < t->n; ++i) {
        const EnvState& e = t->e[(size_t)i];
        if (qpos) { qpos[2 * i] = e.q0; qpos[2 * i + 1] = e.q1; }
        if (qvel) { qvel[2 * i] = e.v0; qvel[2 * i + 1] = e.v1; }
        if (target) { target[2 * i] = e.tx; target[2 * i + 1] = e.ty; }
        if (tip) { tip[2 * i] = e.px; tip[2 * i + 1] = e.py; }
        if (step) step[i] = e.step;
        if (episode) episode[i] = e.episode;
        if (qpos_lo) { qpos_lo[2 * i] = e.q0l; qpos_lo[2 * i + 1] = e.q1l; }
    }
}

// same semantics as k_set_state (env.cu): fingertip NULL => forward kinematics of qpos; qpos without qpos_lo => low parts zero
void twin_set_state(void* h, const float* qpos, const float* qvel, const float* target, const float* tip, const int32_t* step,
                    const uint32_t* episode, const float* qpos_lo) {
    Twin* t = (Twin*)h;
    for (int64_t i = 0; i < t->n; ++i) {
        EnvState& e = t->e[(size_t)i];
        if (qpos) { e.q0 = qpos[2 * i]; e.q1 = qpos[2 * i + 1]; e.q0l = e.q1l = 0.f; }
        if (qpos_lo) { e.q0l = qpos_lo[2 * i]; e.q1l = qpos_lo[2 * i + 1]; }
        if (qvel) { e.v0 = qvel[2 * i]; e.v1 = qvel[2 * i + 1]; }
        if (target) { e.tx = target[2 * i]; e.ty = target[2 * i + 1]; }
        if (tip) { e.px = tip[2 * i]; e.py = tip[2 * i + 1]; }
        else if (qpos) fk(e.q0, e.q1, e.px, e.py);
        if (step) e.step = step[i];
        if (episode) e.episode = episode[i];
        refresh_trig(e);
    }
}

}  // extern "C"
