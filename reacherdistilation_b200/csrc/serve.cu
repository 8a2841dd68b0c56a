// Resident env server: the gym surface at batch 1 (BASELINE config 1; /root/reference src/distilation/mlp_train.py:120-139 -- one env, per step
// `pi.act(ob)` then `env.step(ac)`) without a kernel launch or a copy-engine transfer per call.
//
// A launch + two synchronising copies cost ~50 us per host call, 100 x the arithmetic of one env step.  For small host-surface envs
// (n <= 32: one warp, lane = env) the state therefore stays in the registers of ONE resident warp that polls a 64-byte command line in
// mapped host memory; the host call writes {payload, opcode, sequence number} into that line, the warp sees it with its next PCIe read,
// executes (the same step_env / policy_fwd_simt arithmetic as the batch kernels: bit-identical results), posts the reply block into host
// memory and bumps the reply sequence number the host is spinning on.  One round trip is ~5 us.
//   * STEP also evaluates the registered policy (the teacher) on the NEW observation, so the loop `ac = pi.act(ob); ob = env.step(ac)` is one
//     round trip per step (rb_env_act_step_host).
//   * The warp retires by itself after SERVE_IDLE_NS without a command (state back to HBM), so a device-wide synchronize elsewhere in the
//     process waits at most that long; the next host call starts a new one.  Every other entry point that touches the env (device-side
//     step / rollouts / DAgger) retires it first (env_quiesce).
#include <atomic>
#include <chrono>
#include <cstring>

#include <immintrin.h>

#include "common.cuh"
#include "physics.cuh"
#include "policy_simt.cuh"

namespace rb {

constexpr int SERVE_MAX_ENVS = 32;
constexpr unsigned long long SERVE_IDLE_NS = 1000000ull;       // 1 ms
constexpr int REP_W = 17;                                      // reply words per env: obs 11 | reward | done | pdflat 4
enum : uint32_t { OP_RESET = 1, OP_STEP = 2, OP_POLICY = 3, OP_QUIT = 4 };

struct ServeMail {                              // page-locked, mapped host memory
    volatile uint32_t cmd[16];                  // host -> device, ONE 64-byte line: [0] sequence number (written last) [1] opcode [2..13] inline payload (n == 1)
    volatile uint32_t rep_seq;                  // device -> host: sequence number of the last command answered
    volatile uint32_t exit_gen;                 // generation of the last server warp that retired
    uint32_t pad[14];
    float payload[SERVE_MAX_ENVS * 11];         // n > 1: actions [n][2] or observations [n][11]
    float reply[SERVE_MAX_ENVS * REP_W];
};

struct ServeState {
    ServeMail* host = nullptr;
    ServeMail* dev = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t seq = 0, gen = 0;
    float* d_policy = nullptr;
    int nout = 0;
};

__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u32(volatile uint32_t* p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The policy for ONE sample by the whole warp: lane l owns hidden units 2l, 2l + 1.  Every output is the same fmaf chain in the same order as
// policy_fwd_simt (bias first, inputs in ascending order), so the result is bit-identical to the batch kernels' -- in ~1/8 of the time.
template <int NOUT> __device__ __forceinline__ void policy_fwd_warp(const PolicySmem& S, float* hb, const float* ob, int lane, float* pd) {
    float z[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) z[k] = fminf(5.f, fmaxf(-5.f, (ob[k] - S.mu[k]) * S.inv_sd[k]));
    const int j0 = 2 * lane;
    float a0 = S.b1[j0], a1 = S.b1[j0 + 1];
#pragma unroll
    for (int k = 0; k < 11; ++k) { a0 = fmaf(z[k], S.W1[k][j0], a0); a1 = fmaf(z[k], S.W1[k][j0 + 1], a1); }
    __syncwarp();
    hb[j0] = tanhf(a0); hb[j0 + 1] = tanhf(a1);
    __syncwarp();
    float c0 = S.b2[j0], c1 = S.b2[j0 + 1];
#pragma unroll 16
    for (int k = 0; k < HID; ++k) { const float h = hb[k]; c0 = fmaf(h, S.W2[k][j0], c0); c1 = fmaf(h, S.W2[k][j0 + 1], c1); }
    __syncwarp();
    hb[HID + j0] = tanhf(c0); hb[HID + j0 + 1] = tanhf(c1);
    __syncwarp();
    float o = 0.f;
    if (lane < NOUT) {
        o = S.b3[lane];
#pragma unroll 16
        for (int k = 0; k < HID; ++k) o = fmaf(hb[HID + k], S.W3[k][lane], o);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) pd[j] = __shfl_sync(0xffffffffu, o, j);
    if (NOUT == 2) { pd[2] = S.logstd[0]; pd[3] = S.logstd[1]; }
}

template <int NOUT>
__global__ void __launch_bounds__(32) k_env_server(int n, float4* qv, float4* tp, uint4* ctr, uint32_t k0, uint32_t k1, uint32_t offset, const float* policy,
                                                   ServeMail* mail, uint32_t next_seq, uint32_t gen, unsigned long long idle_ns) {
    __shared__ PolicySmem S;
    __shared__ float hb[2 * HID];
    __shared__ float rep[SERVE_MAX_ENVS * REP_W];
    __shared__ float inb[SERVE_MAX_ENVS * 11];
    const int lane = threadIdx.x;
    if (policy) policy_load_smem(S, policy, NOUT);
    __syncwarp();
    EnvState e = lane < n ? load_state(qv, tp, ctr, lane) : zero_state();
    unsigned long long t_idle = gtimer();
    // policy on the observation rows staged in inb (rows [0, n)) -> rep[.][13..16]
    auto eval_policy = [&]() {
        if (n <= 4) {
            for (int i = 0; i < n; ++i) {
                float ob[11], pd[4];
#pragma unroll
                for (int k = 0; k < 11; ++k) ob[k] = inb[i * 11 + k];
                policy_fwd_warp<NOUT>(S, hb, ob, lane, pd);
                if (lane < 4) rep[i * REP_W + 13 + lane] = pd[lane];
            }
        } else {
            float ob[11], pd[4];
#pragma unroll
            for (int k = 0; k < 11; ++k) ob[k] = lane < n ? inb[lane * 11 + k] : 0.f;
            policy_fwd_simt<NOUT>(S, ob, pd);
            if (lane < n) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rep[lane * REP_W + 13 + j] = pd[j];
            }
        }
        __syncwarp();
    };
    for (;;) {
        const uint32_t w = lane < 16 ? ld_sys_u32(mail->cmd + lane) : 0u;      // the whole command line in one 64-byte read
        const uint32_t seq = __shfl_sync(0xffffffffu, w, 0);
        if (seq != next_seq) {
            const bool expired = gtimer() - t_idle > idle_ns;
            if (__shfl_sync(0xffffffffu, expired ? 1 : 0, 0)) break;
            continue;
        }
        const uint32_t op = __shfl_sync(0xffffffffu, w, 1);
        if (op == OP_QUIT) { next_seq += 1u; break; }
        // ---- payload: inline words 2..13 when n == 1, else the payload area (a second read, ordered behind the sequence number) ----------
        const int pw = op == OP_STEP ? 2 * n : (op == OP_POLICY ? 11 * n : 0);
        if (n == 1) {
            const float v = __uint_as_float(__shfl_sync(0xffffffffu, w, (lane + 2) & 31));
            if (lane < pw) inb[lane] = v;
        } else if (pw > 0) {
            __threadfence_system();
            for (int i = lane; i < pw; i += 32) inb[i] = ld_sys_f32(mail->payload + i);
        }
        __syncwarp();
        if (op == OP_RESET || op == OP_STEP) {
            float ob[11];
            if (lane < n) {
                if (op == OP_RESET) {
                    e.episode = 0u;
                    reset_env(e, k0, k1, offset + (uint32_t)lane);
                    rep[lane * REP_W + 11] = 0.f; rep[lane * REP_W + 12] = 0.f;
                } else {
                    const float a0 = inb[2 * lane], a1 = inb[2 * lane + 1];
                    bool d;
                    const float r = step_env(e, a0, a1, k0, k1, offset + (uint32_t)lane, d);
                    rep[lane * REP_W + 11] = r;
                    rep[lane * REP_W + 12] = __uint_as_float(d ? 1u : 0u);
                }
                observe(e, ob);
            }
            __syncwarp();
            if (lane < n) {
#pragma unroll
                for (int k = 0; k < 11; ++k) { rep[lane * REP_W + k] = ob[k]; inb[lane * 11 + k] = ob[k]; }
            }
            __syncwarp();
            if (policy) eval_policy();                       // the teacher's answer for the observation the caller is about to receive
        } else if (op == OP_POLICY) {
            if (policy) eval_policy();
        }
        __syncwarp();
        for (int i = lane; i < n * REP_W; i += 32) mail->reply[i] = rep[i];       // coalesced posted writes into host memory
        __threadfence_system();
        __syncwarp();
        if (lane == 0) st_sys_u32(&mail->rep_seq, next_seq);
        next_seq += 1u;
        t_idle = gtimer();
    }
    if (lane < n) store_state(qv, tp, ctr, lane, e);
    __threadfence_system();
    __syncwarp();
    if (lane == 0) { st_sys_u32(&mail->rep_seq, next_seq - 1u); st_sys_u32(&mail->exit_gen, gen); }
}

static ServeState* serve_state(rb_env* e) { return reinterpret_cast<ServeState*>(e->serve); }

static int serve_init(rb_env* e) {
    if (e->serve) return RB_OK;
    DeviceGuard guard(e->device);
    ServeState* s = new ServeState();
    void* h = nullptr;
    cudaError_t err = cudaHostAlloc(&h, sizeof(ServeMail), cudaHostAllocMapped);
    if (err == cudaSuccess) { memset(h, 0, sizeof(ServeMail)); err = cudaHostGetDevicePointer((void**)&s->dev, h, 0); }
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&s->d_policy, sizeof(float) * rb_policy_param_count(4));
    if (err != cudaSuccess) { if (h) cudaFreeHost(h); if (s->stream) cudaStreamDestroy(s->stream); delete s; return cuda_fail(err, "serve_init"); }
    s->host = (ServeMail*)h;
    e->serve = s;
    return RB_OK;
}

static bool serve_running(const ServeState* s) { return s->gen != 0 && s->host->exit_gen != s->gen; }

static int serve_launch(rb_env* e) {
    ServeState* s = serve_state(e);
    DeviceGuard guard(e->device);
    s->gen += 1u;
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    const float* pol = s->nout ? s->d_policy : nullptr;
    if (s->nout == 4) k_env_server<4><<<1, 32, 0, s->stream>>>((int)e->n, e->qv, e->tp, e->ctr, k0, k1, e->offset, pol, s->dev, s->seq + 1u, s->gen, SERVE_IDLE_NS);
    else k_env_server<2><<<1, 32, 0, s->stream>>>((int)e->n, e->qv, e->tp, e->ctr, k0, k1, e->offset, pol, s->dev, s->seq + 1u, s->gen, SERVE_IDLE_NS);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

// post {payload, opcode, sequence} and wait for the reply; (re)starts the server warp when none is resident
static int serve_call(rb_env* e, uint32_t op, const float* payload, int words) {
    ServeState* s = serve_state(e);
    ServeMail* m = s->host;
    if (!serve_running(s)) { int rc = serve_launch(e); if (rc) return rc; }
    const uint32_t seq = s->seq + 1u;
    if (e->n == 1) { for (int i = 0; i < words; ++i) { uint32_t u; memcpy(&u, payload + i, 4); m->cmd[2 + i] = u; } }
    else if (words) memcpy(m->payload, payload, sizeof(float) * words);
    m->cmd[1] = op;
    std::atomic_signal_fence(std::memory_order_seq_cst);        // x86 keeps the store order; this keeps the compiler from changing it
    _mm_sfence();
    m->cmd[0] = seq;                                            // published last
    s->seq = seq;
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t spins = 0;
    while (m->rep_seq != seq) {
        _mm_pause();
        if ((++spins & 0xFFFu) == 0) {
            if (!serve_running(s) && m->rep_seq != seq) {       // the warp retired (idle time-out) just before this command was posted: a new one picks it up
                s->seq = seq - 1u;
                int rc = serve_launch(e);
                s->seq = seq;
                if (rc) return rc;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                const cudaError_t q = cudaStreamQuery(s->stream);
                set_error("env server: no reply within 20 s (stream status %d: %s)", (int)q, cudaGetErrorString(q));
                return RB_ERR_CUDA;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return RB_OK;
}

// retire the server warp (state back in HBM); no-op when none is resident
int env_quiesce(rb_env* e) {
    if (!e || !e->serve) return RB_OK;
    ServeState* s = serve_state(e);
    if (serve_running(s)) {
        ServeMail* m = s->host;
        const uint32_t seq = s->seq + 1u;
        m->cmd[1] = OP_QUIT;
        _mm_sfence();
        m->cmd[0] = seq;
        s->seq = seq;
    }
    if (s->gen) { DeviceGuard guard(e->device); RB_CUDA(cudaStreamSynchronize(s->stream)); }
    return RB_OK;
}

void env_serve_destroy(rb_env* e) {
    if (!e || !e->serve) return;
    env_quiesce(e);
    ServeState* s = serve_state(e);
    DeviceGuard guard(e->device);
    cudaFree(s->d_policy);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->host) cudaFreeHost(s->host);
    delete s;
    e->serve = nullptr;
}

bool env_serve_eligible(const rb_env* e) { return e->n <= SERVE_MAX_ENVS; }

static void unpack_reply(const rb_env* e, float* obs, float* rew, uint8_t* done, float* pd) {
    const ServeMail* m = serve_state(const_cast<rb_env*>(e))->host;
    for (int i = 0; i < (int)e->n; ++i) {
        const float* r = m->reply + i * REP_W;
        if (obs) memcpy(obs + i * 11, r, sizeof(float) * 11);
        if (rew) rew[i] = r[11];
        if (done) { uint32_t u; memcpy(&u, r + 12, 4); done[i] = (uint8_t)u; }
        if (pd) memcpy(pd + i * 4, r + 13, sizeof(float) * 4);
    }
}

int env_serve_reset(rb_env* e, float* obs_host, float* pd_host) {
    int rc = serve_init(e);
    if (rc) return rc;
    RB_REQUIRE(pd_host == nullptr || serve_state(e)->nout != 0, "no policy registered (rb_env_serve_policy)");
    rc = serve_call(e, OP_RESET, nullptr, 0);
    if (rc) return rc;
    unpack_reply(e, obs_host, nullptr, nullptr, pd_host);
    return RB_OK;
}

int env_serve_step(rb_env* e, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* pd_host) {
    int rc = serve_init(e);
    if (rc) return rc;
    RB_REQUIRE(pd_host == nullptr || serve_state(e)->nout != 0, "no policy registered (rb_env_serve_policy)");
    rc = serve_call(e, OP_STEP, act_host, 2 * (int)e->n);
    if (rc) return rc;
    unpack_reply(e, obs_host, rew_host, done_host, pd_host);
    return RB_OK;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_env_serve_policy(rb_env* e, const float* params_host, int nout) {
    RB_REQUIRE(e != nullptr && params_host != nullptr, "NULL argument");
    RB_REQUIRE(nout == 2 || nout == 4, "nout must be 2 or 4");
    RB_REQUIRE(env_serve_eligible(e), "the resident env server takes at most 32 envs");
    int rc = serve_init(e);
    if (rc) return rc;
    rc = env_quiesce(e);                                        // the resident warp holds the previous policy in shared memory
    if (rc) return rc;
    ServeState* s = serve_state(e);
    DeviceGuard guard(e->device);
    RB_CUDA(cudaMemcpyAsync(s->d_policy, params_host, sizeof(float) * rb_policy_param_count(nout), cudaMemcpyHostToDevice, s->stream));
    RB_CUDA(cudaStreamSynchronize(s->stream));
    s->nout = nout;
    return RB_OK;
}

int rb_env_serve_policy_fwd(rb_env* e, const float* obs_host, float* pd_host) {
    RB_REQUIRE(e != nullptr && obs_host != nullptr && pd_host != nullptr, "NULL argument");
    RB_REQUIRE(e->serve != nullptr && serve_state(e)->nout != 0, "no policy registered (rb_env_serve_policy)");
    int rc = serve_call(e, OP_POLICY, obs_host, 11 * (int)e->n);
    if (rc) return rc;
    unpack_reply(e, nullptr, nullptr, nullptr, pd_host);
    return RB_OK;
}

int rb_env_act_step_host(rb_env* e, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* pd_next_host) {
    RB_REQUIRE(e != nullptr && act_host != nullptr && obs_host != nullptr, "NULL argument");
    RB_REQUIRE(env_serve_eligible(e), "the resident env server takes at most 32 envs");
    return env_serve_step(e, act_host, obs_host, rew_host, done_host, pd_next_host);
}

}  // extern "C"
