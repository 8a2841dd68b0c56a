// Lock-step DAgger iteration pieces: observe (ob, teacher label, student input) and act (step with the student mean).
// Replaces the per-env-step body of /root/reference src/distilation/mlp_train.py:143-204 (teacher label :165-167, record
// :188-193 with dataset.py:118-143 `prev` / `prew` semantics, env.step(s_ac) :196) for N envs at once.
#include <chrono>
#include <cstring>

#include "common.cuh"
#include "dagger_input.cuh"
#include "physics.cuh"

struct rb_dagger {
    rb_env* env = nullptr;
    // split-weight image of the (frozen) teacher for the tensor-core observe kernel, keyed by the parameter pointer
    void* teacher_img = nullptr;
    const float* teacher_img_src = nullptr;
    int kind = 0;
    float keep_prob = 1.f;
    float4* prev_t = nullptr;      // teacher pdflat of the previous record of the current episode
    float* prev_rec_rew = nullptr; // 'rew' field of the previous record
    float* last_reward = nullptr;  // reward returned by the last env.step (not cleared at reset, mlp_train.py:114,196)
    // rb_dagger_step: device-side step clock {iteration, adam step, exchange epoch} and the captured CUDA graph of one iteration
    uint32_t* clock = nullptr;
    cudaGraphExec_t gexec = nullptr;
    uint64_t gkey = 0;
    cudaStream_t cap_stream = nullptr;   // capture happens here (the legacy default stream cannot be captured); replay on the caller's stream
    cudaStream_t side_stream = nullptr;  // forked branch: the student's weight image is built beside the observe kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // result mailbox in page-locked mapped HOST memory: {loss bits, iteration count} stored by the last kernel of rb_dagger_step as one 8-byte
    // word; rb_dagger_wait_loss polls it, so the per-iteration loss read-back (mlp_train.py:148-161 prints it) needs no stream synchronise
    // and no copy-engine transfer
    volatile uint2* mailbox_host = nullptr;
    uint2* mailbox_dev = nullptr;
    uint32_t* act_flags = nullptr;       // per-tile forward-done flags of the env step fused into k_student_tc (student_tc.cu: act_*)
};

namespace rb {

__global__ void k_dagger_input(int64_t n, const uint4* __restrict__ ctr, const float* __restrict__ obs, const float4* __restrict__ prev_t,
                               const float* __restrict__ prev_rec_rew, float keep_prob, uint32_t k0, uint32_t k1, uint32_t offset,
                               uint32_t iteration, float4* __restrict__ x, float4* __restrict__ x_act) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ob[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) ob[k] = __ldg(obs + i * 11 + k);
    const bool first = ctr[i].x == 0u;   // first record of an episode: prev / prew are zeros (dataset.py:151-164)
    const float4 pp = first ? make_float4(0.f, 0.f, 0.f, 0.f) : prev_t[i];
    const float pr = first ? 0.f : prev_rec_rew[i];
    float4 o[4];
    mlp_input_row(ob, keep_prob, k0, k1, offset + (uint32_t)i, iteration, pp, pr, o);
    x[i * 4 + 0] = o[0]; x[i * 4 + 1] = o[1]; x[i * 4 + 2] = o[2]; x[i * 4 + 3] = o[3];
    if (x_act) {                         // the same row without observation dropout: what the student ACTS on (mlp_train.py:171-186, keep_prob 1)
        x_act[i * 4 + 0] = make_float4(ob[0], ob[1], ob[2], ob[3]); x_act[i * 4 + 1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
        x_act[i * 4 + 2] = make_float4(ob[8], ob[9], ob[10], pp.x); x_act[i * 4 + 3] = make_float4(pp.y, pp.z, pp.w, pr);
    }
}

__global__ void __launch_bounds__(128) k_dagger_act(int64_t n, float4* qv, float4* tp, uint4* ctr, const float4* __restrict__ s_pd,
                                                    const float4* __restrict__ t_pd, float4* prev_t, float* prev_rec_rew, float* last_reward,
                                                    float* __restrict__ rew, uint8_t* __restrict__ done, uint32_t k0, uint32_t k1,
                                                    uint32_t offset, uint32_t* clock, const float* __restrict__ loss_src, uint2* mailbox, int64_t i0) {
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // envs [i0, n)
    if (clock && i == i0) {                                                       // end of the iteration: advance the device-side step clock
        const uint32_t it = clock[0] + 1u;
        clock[0] = it; clock[1] += 1u; clock[2] += 1u;
        if (mailbox)                                                              // {loss, iterations done} -> host memory, one 8-byte posted write
            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(mailbox), "r"(__float_as_uint(*loss_src)), "r"(it) : "memory");
    }
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

__global__ void k_set_clock(uint32_t* clock, uint32_t iteration, uint32_t adam_t, uint32_t epoch) {
    clock[0] = iteration; clock[1] = adam_t; clock[2] = epoch; clock[3] = 0u;
}

// implemented in student_tc.cu
struct AdamFuse { float* p; float* m; float* v; float lr_t, beta1, beta2, eps, gscale; };
struct PeerExchange { int world, rank; uint32_t epoch; const uint64_t* gl_ptrs; const uint64_t* flag_ptrs; const uint64_t* gl_ptrs_alt; };
struct StepClock { const uint32_t* clock; float lr, beta1, beta2; };
struct ActFuse {
    float4* qv; float4* tp; uint4* ctr; float4* prev_t; float* prev_rec_rew; float* last_reward; float* rew; uint8_t* done;
    uint32_t k0, k1, offset; uint32_t* flags; uint32_t* clock; uint2* mailbox; const float* x_act; int image_prebuilt;
};
int student_tc_run_ex(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only, float* s_out,
                      float* gradloss, void* workspace, const AdamFuse* adam, const PeerExchange* px, const StepClock* clk, const ActFuse* act,
                      cudaStream_t st);
int64_t student_tc_act_covered(int64_t B, int grid);
int student_tc_build_image(int kind, const float* params, void* workspace, const StepClock* clk, cudaStream_t st);
int student_tc_grid(int* grid);

// implemented in policy_tc.cu
size_t policy_tc_image_bytes();
int policy_tc_build_image(const float* params, int nout, void* img, cudaStream_t s);
int dagger_observe_tc(rb_env* e, const void* teacher_img, int student_kind, float keep_prob, const float4* prev_t, const float* prev_rec_rew,
                      uint32_t iteration, const uint32_t* clock, float* obs, float* t_pd, float* x, float* x_act, cudaStream_t s);

}  // namespace rb

using namespace rb;

extern "C" {

int rb_dagger_create(rb_dagger** out, rb_env* env, int student_kind, float keep_prob) {
    RB_REQUIRE(out && env, "NULL argument");
    RB_REQUIRE(student_kind == RB_STUDENT_POLICY64 || student_kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(keep_prob > 0.f, "keep_prob must be > 0");
    DeviceGuard guard(env->device);
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
    cudaFree(d->prev_t); cudaFree(d->prev_rec_rew); cudaFree(d->last_reward); cudaFree(d->teacher_img); cudaFree(d->clock); cudaFree(d->act_flags);
    if (d->mailbox_host) cudaFreeHost((void*)d->mailbox_host);
    if (d->gexec) cudaGraphExecDestroy(d->gexec);
    if (d->cap_stream) cudaStreamDestroy(d->cap_stream);
    if (d->side_stream) cudaStreamDestroy(d->side_stream);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->ev_join) cudaEventDestroy(d->ev_join);
    delete d;
    return RB_OK;
}

// Checkpoint / resume (`train(train, restore)`, lstm_train.py:86-87,102-107): the per-env loop state the handle carries from one
// iteration to the next.  Plain device-to-device copies on the caller's stream; a NULL pointer skips that field.
int rb_dagger_get_state(rb_dagger* d, float* prev_t, float* prev_rec_rew, float* last_reward, void* stream) {
    RB_REQUIRE(d != nullptr, "dagger handle is NULL");
    const size_t n = (size_t)d->env->n;
    cudaStream_t s = (cudaStream_t)stream;
    if (prev_t) RB_CUDA(cudaMemcpyAsync(prev_t, d->prev_t, sizeof(float4) * n, cudaMemcpyDeviceToDevice, s));
    if (prev_rec_rew) RB_CUDA(cudaMemcpyAsync(prev_rec_rew, d->prev_rec_rew, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    if (last_reward) RB_CUDA(cudaMemcpyAsync(last_reward, d->last_reward, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    return RB_OK;
}

int rb_dagger_set_state(rb_dagger* d, const float* prev_t, const float* prev_rec_rew, const float* last_reward, void* stream) {
    RB_REQUIRE(d != nullptr, "dagger handle is NULL");
    const size_t n = (size_t)d->env->n;
    cudaStream_t s = (cudaStream_t)stream;
    if (prev_t) RB_CUDA(cudaMemcpyAsync(d->prev_t, prev_t, sizeof(float4) * n, cudaMemcpyDeviceToDevice, s));
    if (prev_rec_rew) RB_CUDA(cudaMemcpyAsync(d->prev_rec_rew, prev_rec_rew, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    if (last_reward) RB_CUDA(cudaMemcpyAsync(d->last_reward, last_reward, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    return RB_OK;
}

int rb_dagger_observe(rb_dagger* d, const float* teacher_params, uint32_t iteration, float* obs, float* t_pd, float* x, float* x_act, int mode,
                      void* stream) {
    RB_REQUIRE(d && teacher_params && obs && t_pd && x, "NULL argument");
    rb_env* e = d->env;
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    if (mode == RB_MODE_TC) {            // one fused kernel: observe + teacher (tcgen05) + student input
        if (!d->teacher_img) RB_CUDA(cudaMalloc(&d->teacher_img, policy_tc_image_bytes()));
        if (d->teacher_img_src != teacher_params) {
            int rc0 = policy_tc_build_image(teacher_params, 2, d->teacher_img, (cudaStream_t)stream);
            if (rc0) return rc0;
            d->teacher_img_src = teacher_params;
        }
        return dagger_observe_tc(e, d->teacher_img, d->kind, d->keep_prob, (const float4*)d->prev_t, d->prev_rec_rew, iteration, nullptr, obs, t_pd, x,
                                 d->kind == RB_STUDENT_MLP ? x_act : nullptr, (cudaStream_t)stream);
    }
    int rc = rb_env_observe(e, obs, stream);
    if (rc) return rc;
    rc = rb_policy_fwd(teacher_params, 2, obs, e->n, t_pd, mode, stream);
    if (rc) return rc;
    if (d->kind == RB_STUDENT_POLICY64) {
        if (x != obs) RB_CUDA(cudaMemcpyAsync(x, obs, sizeof(float) * OBS * e->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    } else {
        k_dagger_input<<<(unsigned)((e->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->n, e->ctr, obs, d->prev_t, d->prev_rec_rew, d->keep_prob,
                                                                                        (uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->offset,
                                                                                        iteration, (float4*)x, (float4*)x_act);
        RB_CUDA(cudaGetLastError());
    }
    return RB_OK;
}

int rb_dagger_invalidate_teacher(rb_dagger* d) {
    RB_REQUIRE(d != nullptr, "NULL argument");
    d->teacher_img_src = nullptr;
    return RB_OK;
}

int rb_dagger_act(rb_dagger* d, const float* s_pd, const float* t_pd, float* rew, uint8_t* done, void* stream) {
    RB_REQUIRE(d && s_pd, "NULL argument");
    rb_env* e = d->env;
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    k_dagger_act<<<(unsigned)((e->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e->n, e->qv, e->tp, e->ctr, (const float4*)s_pd, (const float4*)t_pd,
                                                                                  d->prev_t, d->prev_rec_rew, d->last_reward, rew, done,
                                                                                  (uint32_t)e->seed, (uint32_t)(e->seed >> 32), e->offset, nullptr, nullptr, nullptr, 0);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_dagger_set_clock(rb_dagger* d, uint32_t iteration, uint32_t adam_t, uint32_t epoch, void* stream) {
    RB_REQUIRE(d != nullptr, "NULL argument");
    if (!d->clock) RB_CUDA(cudaMalloc(&d->clock, 4 * sizeof(uint32_t)));
    const size_t nflags = (size_t)((d->env->n + 127) / 128);
    if (!d->act_flags) RB_CUDA(cudaMalloc(&d->act_flags, nflags * sizeof(uint32_t)));
    RB_CUDA(cudaMemsetAsync(d->act_flags, 0, nflags * sizeof(uint32_t), (cudaStream_t)stream));      // flag values are iteration counts: a clock reset voids them
    k_set_clock<<<1, 1, 0, (cudaStream_t)stream>>>(d->clock, iteration, adam_t, epoch);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

// One whole DAgger iteration (observe + teacher label + student input | student loss, gradient [, peer all-reduce], Adam | env step) as
// three launches that take every per-step quantity from the device-side clock, so the sequence is captured once in a CUDA graph and
// replayed with a single cudaGraphLaunch.
int rb_dagger_step(rb_dagger* d, const float* teacher_params, float* params, float* m, float* v, float* gradloss, void* ws, float* obs, float* t_pd,
                   float* x, float* x_act, float* s_pd, float* rew, uint8_t* done, int loss_kind, float lr, float b1, float b2, float eps, float gscale, int rank,
                   int world, const uint64_t* slots_even, const uint64_t* slots_odd, const uint64_t* flags, int use_graph, void* stream) {
    RB_REQUIRE(d && teacher_params && params && m && v && gradloss && ws && obs && t_pd && x && s_pd, "NULL argument");
    RB_REQUIRE(d->clock != nullptr, "call rb_dagger_set_clock first");
    RB_REQUIRE(world == 1 || (world >= 2 && world <= 8 && slots_even && slots_odd && flags), "data parallel: 2..8 ranks with peer slots and flags");
    rb_env* e = d->env;
    { const int qrc = env_quiesce(e); if (qrc) return qrc; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t loss_index = rb_student_param_count(d->kind);          // gradloss = [grad P | loss]
    if (!d->mailbox_host) {
        void* h = nullptr;
        RB_CUDA(cudaHostAlloc(&h, sizeof(uint2), cudaHostAllocMapped));
        memset(h, 0, sizeof(uint2));
        void* dv = nullptr;
        RB_CUDA(cudaHostGetDevicePointer(&dv, h, 0));
        d->mailbox_host = (volatile uint2*)h; d->mailbox_dev = (uint2*)dv;
    }
    if (!d->teacher_img) RB_CUDA(cudaMalloc(&d->teacher_img, policy_tc_image_bytes()));
    if (d->teacher_img_src != teacher_params) {
        int rc0 = policy_tc_build_image(teacher_params, 2, d->teacher_img, st);
        if (rc0) return rc0;
        d->teacher_img_src = teacher_params;
    }
    constexpr bool fuse_act = true;      // the env step runs inside the student launch (the three-launch form below is kept for reference timing)
    if (!d->side_stream) {
        RB_CUDA(cudaStreamCreateWithFlags(&d->side_stream, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
    }
    auto issue = [&](cudaStream_t st) -> int {
        const float* xa = d->kind == RB_STUDENT_MLP ? x_act : nullptr;       // the 2x64 student sees the raw observation: nothing to un-drop
        // fork: the student's weight image (bf16 hi / lo tiles, folded layers, un-fold snapshot) only needs the parameters, so it is built on a
        // second branch beside the observe kernel (two parallel nodes of the captured graph) and joined in front of the student launch
        RB_CUDA(cudaEventRecord(d->ev_fork, st));
        RB_CUDA(cudaStreamWaitEvent(d->side_stream, d->ev_fork, 0));
        const StepClock clk{d->clock, lr, b1, b2};
        int rc = student_tc_build_image(d->kind, params, ws, &clk, d->side_stream);
        if (rc) return rc;
        RB_CUDA(cudaEventRecord(d->ev_join, d->side_stream));
        rc = dagger_observe_tc(e, d->teacher_img, d->kind, d->keep_prob, (const float4*)d->prev_t, d->prev_rec_rew, 0u, d->clock, obs, t_pd, x, (float*)xa, st);
        if (rc) return rc;
        RB_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));                      // join
        const AdamFuse af{params, m, v, 0.f, b1, b2, eps, gscale};
        const PeerExchange px{world, rank, 0u, slots_even, flags, slots_odd};
        if (fuse_act) {      // env step, clock advance and loss mailbox inside the student launch: two launches per iteration
            const ActFuse act{e->qv, e->tp, e->ctr, d->prev_t, d->prev_rec_rew, d->last_reward, rew, done, (uint32_t)e->seed, (uint32_t)(e->seed >> 32),
                              e->offset, d->act_flags, d->clock, d->mailbox_dev, xa, 1};
            rc = student_tc_run_ex(d->kind, params, x, t_pd, e->n, loss_kind, 0, s_pd, gradloss, ws, &af, world > 1 ? &px : nullptr, &clk, &act, st);
            if (rc) return rc;
            int grid = 1;
            rc = student_tc_grid(&grid);
            if (rc) return rc;
            const int64_t i0 = student_tc_act_covered(e->n, grid);           // envs the student launch did not step (none at the config-4 shard)
            if (i0 < e->n) {
                k_dagger_act<<<(unsigned)((e->n - i0 + 127) / 128), 128, 0, st>>>(e->n, e->qv, e->tp, e->ctr, (const float4*)s_pd, (const float4*)t_pd, d->prev_t,
                                                                                   d->prev_rec_rew, d->last_reward, rew, done, (uint32_t)e->seed,
                                                                                   (uint32_t)(e->seed >> 32), e->offset, nullptr, nullptr, nullptr, i0);
                RB_CUDA(cudaGetLastError());
            }
            return RB_OK;
        }
        RB_REQUIRE(xa == nullptr, "x_act needs the env step fused into the student launch");
        rc = student_tc_run_ex(d->kind, params, x, t_pd, e->n, loss_kind, 0, s_pd, gradloss, ws, &af, world > 1 ? &px : nullptr, &clk, nullptr, st);
        if (rc) return rc;
        k_dagger_act<<<(unsigned)((e->n + 127) / 128), 128, 0, st>>>(e->n, e->qv, e->tp, e->ctr, (const float4*)s_pd, (const float4*)t_pd, d->prev_t,
                                                                      d->prev_rec_rew, d->last_reward, rew, done, (uint32_t)e->seed,
                                                                      (uint32_t)(e->seed >> 32), e->offset, d->clock, gradloss + loss_index, d->mailbox_dev, 0);
        RB_CUDA(cudaGetLastError());
        return RB_OK;
    };
    if (!use_graph) return issue(st);
    // graph key: everything the captured launches bake in
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t vv) { key = (key ^ vv) * 1099511628211ull; };
    const void* ptrs[] = {teacher_params, params, m, v, gradloss, ws, obs, t_pd, x, x_act, s_pd, rew, done, stream, slots_even, slots_odd, flags};
    for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
    const float fl[] = {lr, b1, b2, eps, gscale};
    for (float f : fl) { uint32_t u; memcpy(&u, &f, 4); mix(u); }
    mix((uint64_t)loss_kind); mix((uint64_t)rank); mix((uint64_t)world);
    if (world > 1) for (int r = 0; r < world; ++r) { mix(slots_even[r]); mix(slots_odd[r]); mix(flags[r]); }
    if (!d->gexec || d->gkey != key) {
        if (d->gexec) { cudaGraphExecDestroy(d->gexec); d->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        if (!d->cap_stream) RB_CUDA(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
        RB_CUDA(cudaStreamSynchronize(st));                            // e.g. the teacher image build above
        RB_CUDA(cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeRelaxed));
        const int rc = issue(d->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(d->cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        const cudaError_t ci = cudaGraphInstantiate(&d->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) { d->gexec = nullptr; return cuda_fail(ci, "cudaGraphInstantiate"); }
        d->gkey = key;
    }
    RB_CUDA(cudaGraphLaunch(d->gexec, st));
    return RB_OK;
}

// Blocks until the rb_dagger_step iteration that brought the device clock's iteration count to `iteration` has finished and returns its
// loss (summed over ranks in the data-parallel case).  Polls the mapped host mailbox: no stream synchronise, no copy.
int rb_dagger_wait_loss(rb_dagger* d, uint32_t iteration, float* loss_out) {
    RB_REQUIRE(d && loss_out, "NULL argument");
    RB_REQUIRE(d->mailbox_host != nullptr, "rb_dagger_step has not run yet");
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t spins = 0;
    for (;;) {
        const uint64_t w = *reinterpret_cast<volatile uint64_t*>(d->mailbox_host);      // one aligned 8-byte read: {loss, iteration} are consistent
        const uint32_t it = (uint32_t)(w >> 32);
        if ((int32_t)(it - iteration) >= 0) {
            RB_REQUIRE(it == iteration, "rb_dagger_wait_loss: that iteration has already been overwritten by a later one");
            const uint32_t bits = (uint32_t)w;
            memcpy(loss_out, &bits, 4);
            return RB_OK;
        }
        if ((++spins & 0xFFFu) == 0) {
            const cudaError_t err = cudaPeekAtLastError();
            if (err != cudaSuccess) return cuda_fail(err, "rb_dagger_wait_loss");
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) { set_error("rb_dagger_wait_loss: timed out"); return RB_ERR_CUDA; }
        }
    }
}

}  // extern "C"
