/* reacher_b200.h -- C ABI of libreacher_b200.so: B200 (sm_100a) hot path of ReacherDistilation.
 *
 * The reference (winstonww/ReacherDistilation, paths below relative to its root) has no FFI layer: its seam is the
 * gym env protocol plus the TF session calls inside the training loops.  Each entry point here replaces one of those
 * call sites; the Python host layer (reacherdistilation_b200/) binds them with ctypes and exposes the reference's names.
 *
 * Conventions
 *   - every function returns 0 on success, a negative rb_status on failure; rb_last_error() gives the message
 *     (thread-local).  Nothing here falls back to the CPU: without a CUDA device every call fails with RB_ERR_CUDA.
 *   - pointers named *_dev are DEVICE pointers (the caller owns them, e.g. torch tensors); functions taking them are
 *     ASYNCHRONOUS on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream), perform no hidden
 *     synchronisation and no host round trip.  Functions suffixed _host take HOST pointers, copy in/out and synchronise.
 *   - a handle is bound to one GPU and is not thread-safe; use one handle per rank.
 *   - all floating-point data is IEEE fp32 unless stated; obs rows are 11 floats, pdflat rows 4 floats
 *     (mean0, mean1, logstd0, logstd1), actions 2 floats; row-major, env index outermost ([N,11] ...),
 *     rollout buffers time-major ([T,N,11] ...).
 */
#ifndef REACHER_B200_H
#define REACHER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RB_OK = 0,
    RB_ERR_INVALID = -1, /* bad argument */
    RB_ERR_CUDA = -2,    /* CUDA runtime error / no device */
    RB_ERR_ALLOC = -3,
    RB_ERR_UNSUPPORTED = -4
} rb_status;

#define RB_OBS_DIM 11
#define RB_ACT_DIM 2
#define RB_PDFLAT_DIM 4
#define RB_EPISODE_STEPS 50

/* compute mode of the dense layers */
#define RB_MODE_FP32 0 /* CUDA-core fp32 FMA (strict parity path)                         */
#define RB_MODE_TC 1   /* tcgen05 tensor cores, bf16x3 split operands, fp32 accumulate in TMEM */

/* student kinds */
#define RB_STUDENT_POLICY64 0 /* backup/student_rollout.py:79-87  MlpPolicy 11-64-64-4, obfilter clip +-5      */
#define RB_STUDENT_MLP 1      /* student_nn.py:51-57  16-24-128-128(lin)-32-4 on [dropout(ob), prev_pdflat, prev_rew] */

/* loss kinds */
#define RB_LOSS_KL_ST 0 /* loss.py:3-13  KL(student || teacher), summed                       */
#define RB_LOSS_KL_TS 1 /* backup/student_rollout.py:639-640  KL(teacher || student), summed  */
#define RB_LOSS_MSE 2   /* backup/student_rollout_mlp_vf.py:276, student_rollout.py:328  sum of squared errors over every output
                           (students: over the 4 pdflat entries)                                                          */
#define RB_LOSS_MSE_ACTION 3 /* students only: sum of squared errors on the MEAN half of pdflat, i.e. on the actions the two
                                policies would take (mlp_train.py:63-66 acts with the mean); the logstd outputs get no gradient */

const char* rb_last_error(void);
int rb_version(void);
/* number of CUDA devices visible (<=0: none) and SM count of `device` */
int rb_device_count(void);
int rb_sm_count(int device);
/* 1 if the given RB_MODE_* is compiled into this build (RB_MODE_TC needs the tcgen05 kernels) */
int rb_mode_available(int mode);          /* policy / rollout kernels */
int rb_student_mode_available(int mode);  /* student forward/backward kernels */

/* ------------------------------------------------------------------------------------------------ env ----
 * rb_env = N lock-step Reacher-v2 environments resident in HBM.
 * Replaces make_mujoco_env("Reacher-v2", 0)           src/distilation/mlp_train.py:21, lstm_train.py:21, teacher.py:29,40
 * global_env_offset: first global env id of this shard -- Philox streams are keyed by GLOBAL env id so any
 * partition of the envs over GPUs reproduces the same per-env trajectories.                                   */
typedef struct rb_env rb_env;
int rb_env_create(rb_env** out, int64_t num_envs, uint64_t seed, int device, uint32_t global_env_offset);
int rb_env_destroy(rb_env* env);
int64_t rb_env_num_envs(const rb_env* env);

/* env.reset()  src/distilation/mlp_train.py:112 -- all envs to episode 0 of their Philox stream; obs_dev[N,11] */
int rb_env_reset(rb_env* env, float* obs_dev, void* stream);
/* env.step(a)  src/distilation/mlp_train.py:135,196 ; lstm_train.py:133,192
 * act_dev[N,2] -> obs_dev[N,11], rew_dev[N], done_dev[N] (1 at the 50th step).  Finished envs are auto-reset and
 * return their reset observation (the reference caller does `if new: ob = env.reset()`, mlp_train.py:137-139).   */
int rb_env_step(rb_env* env, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* done_dev, void* stream);
/* same calls with HOST buffers -- the gym-style surface for a host-resident caller.  N > 32: copies + launch + synchronise inside the call.
 * N <= 32 (the reference's own shape, ONE env: BASELINE config 1): a resident server warp keeps the env state in registers and is driven
 * through a 64-byte command line in mapped host memory -- no launch, no copy-engine transfer, ~5 us per call (csrc/serve.cu); it retires by
 * itself after 1 ms without a call, and any device-side entry point on the same env retires it first.  Same arithmetic: bit-identical. */
int rb_env_reset_host(rb_env* env, float* obs_host);
int rb_env_step_host(rb_env* env, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host);
/* The per-step loop of the teacher warm-up at batch 1  src/distilation/mlp_train.py:120-139  (`ac = pi.act(ob)`; `ob, r, new, _ = env.step(ac)`)
 * on the resident server (N <= 32): rb_env_serve_policy registers the policy (teacher.py:12-20; flat layout of rb_policy_param_count);
 * rb_env_act_step_host = env.step(act) AND the policy's pdflat for the observation it returns (pd_next_host[N,4], may be NULL), one round
 * trip; rb_env_serve_policy_fwd = sess.run(pi.pd.flat) for arbitrary observation rows obs_host[N,11].  Results are bit-identical to
 * rb_env_step / rb_policy_fwd(RB_MODE_FP32).                                                                                        */
int rb_env_serve_policy(rb_env* env, const float* params_host, int nout);
int rb_env_serve_policy_fwd(rb_env* env, const float* obs_host, float* pdflat_host);
int rb_env_act_step_host(rb_env* env, const float* act_host, float* obs_host, float* rew_host, uint8_t* done_host, float* pd_next_host);

/* explicit state access (parity tests start device and oracle from identical states; checkpoint / resume):
 * qpos_dev[N,2], qvel_dev[N,2], target_dev[N,2], fingertip_dev[N,2] (MuJoCo's stale xpos), step_dev[N] int32,
 * episode_dev[N] uint32, qpos_lo_dev[N,2].  Any pointer may be NULL.  The joint angles are carried as two floats (qpos + qpos_lo,
 * |qpos_lo| <= ulp(qpos) / 2: compensated accumulation over the RK4 sub-steps); qpos alone is the angle rounded to fp32.
 * set: fingertip NULL => recomputed by forward kinematics; qpos given without qpos_lo => low parts zero.                       */
int rb_env_get_state(rb_env* env, float* qpos_dev, float* qvel_dev, float* target_dev, float* fingertip_dev,
                     int32_t* step_dev, uint32_t* episode_dev, float* qpos_lo_dev, void* stream);
int rb_env_set_state(rb_env* env, const float* qpos_dev, const float* qvel_dev, const float* target_dev,
                     const float* fingertip_dev, const int32_t* step_dev, const uint32_t* episode_dev, const float* qpos_lo_dev, void* stream);
/* current observation of every env (no stepping) */
int rb_env_observe(rb_env* env, float* obs_dev, void* stream);

/* Fused T-step rollout with Philox random actions a~U(-1,1)^2 keyed (seed, global env, step0+t)
 * (BASELINE.json config 2).  State stays in registers for all T steps.  Buffers (any may be NULL):
 * obs_buf_dev[T,N,11] = observation AFTER step t, act_buf_dev[T,N,2], rew_buf_dev[T,N], done_buf_dev[T,N].        */
int rb_env_rollout_random(rb_env* env, int T, uint32_t step0, float* obs_buf_dev, float* act_buf_dev, float* rew_buf_dev,
                          uint8_t* done_buf_dev, void* stream);

/* ------------------------------------------------------------------------------------------------ policy -
 * baselines MlpPolicy 11 -> 64 tanh -> 64 tanh -> nout, obfilter z = clip((ob-mean)/std, +-5).
 * params (fp32, flat, rb_policy_param_count(nout) floats):
 *   ob_mean[11] ob_std[11] W1[11][64] b1[64] W2[64][64] b2[64] W3[64][nout] b3[nout] logstd[2]
 * nout = 2: teacher (teacher.py:12-16), pdflat = (mean, logstd).  nout = 4: student a5' (all four from the net). */
int64_t rb_policy_param_count(int nout);
/* sess.run((pi.pd.mean, pi.pd.flat))  src/distilation/mlp_train.py:123-125,165-167 ; obs_dev[n,11] -> pdflat_dev[n,4] */
int rb_policy_fwd(const float* params_dev, int nout, const float* obs_dev, int64_t n, float* pdflat_dev, int mode,
                  void* stream);
int rb_policy_fwd_host(const float* params_host, int nout, const float* obs_host, int64_t n, float* pdflat_host, int mode,
                       int device);

/* Teacher warm-up / replay loop  src/distilation/mlp_train.py:120-139 (and teacher_replay, backup/student_rollout.py:93-120)
 * fused on the device: for t in [0,T): record ob_t, pdflat_t = policy(ob_t); step with the pd mean; record reward, done.
 * Rollout buffer (device resident, any pointer may be NULL):
 *   obs_buf_dev[T,N,11]  observation the policy saw at step t      pd_buf_dev[T,N,4]  its pdflat
 *   rew_buf_dev[T,N]     reward returned by that step              done_buf_dev[T,N]  1 when that step ended the episode */
int rb_env_rollout_policy(rb_env* env, const float* params_dev, int nout, int T, float* obs_buf_dev, float* pd_buf_dev,
                          float* rew_buf_dev, uint8_t* done_buf_dev, int mode, void* stream);
/* host-buffer variant: params_host in, the four buffers out (pinned or pageable host memory), synchronises.
 * The whole rollout buffer is always written on the device (it stays resident for the distillation loop, see
 * rb_env_rollout_buffer); a NULL host pointer only means that field is not brought to the host.  Page-locked (mapped) reward /
 * done buffers are written by the kernel directly; everything else is copied in time slabs that overlap the next slab's kernel. */
int rb_env_rollout_policy_host(rb_env* env, const float* params_host, int nout, int T, float* obs_buf_host, float* pd_buf_host,
                               float* rew_buf_host, uint8_t* done_buf_host, int mode);
/* Same call with two more optional per-env outputs (8 + 4 bytes per env instead of 5 T): done_mask_host[N] uint64, bit t = the env finished
 * an episode (TimeLimit 50) at step t of this call (T <= 64); return_sum_host[N] float = the sum of the env's rewards over the T steps of the
 * call, added in step order (a 50-step call issued at an episode boundary gives the episode return the reference logs, mlp_train.py:129-139).
 * Page-locked buffers are written by the kernel directly: {reward, done_mask} or {return_sum, done_mask} need no copy-engine transfer. */
int rb_env_rollout_policy_host_ex(rb_env* env, const float* params_host, int nout, int T, float* obs_buf_host, float* pd_buf_host,
                                  float* rew_buf_host, uint8_t* done_buf_host, uint64_t* done_mask_host, float* return_sum_host, int mode);
/* Split-phase form for callers that keep the GPU busy: _begin queues {H2D of the parameters, the rollout launch} and returns, _wait blocks until
 * the OLDEST outstanding call has finished (its outputs are then complete in host memory).  Up to two calls may be in flight, so the host-side
 * work around call i overlaps the kernel of call i + 1.  Outputs: the kernel-stored ones only (reward [T,N], done_mask [N], return_sum [N]),
 * each NULL or a page-locked buffer that must stay untouched until its _wait.                                                          */
int rb_env_rollout_policy_host_begin(rb_env* env, const float* params_host, int nout, int T, float* rew_buf_host, uint64_t* done_mask_host,
                                     float* return_sum_host, int mode);
int rb_env_rollout_policy_host_wait(rb_env* env);
/* How rb_env_rollout_policy_host brings reward / done into PAGE-LOCKED host buffers: bit 0 set = the kernel stores reward straight into the
 * mapped buffer (posted PCIe writes under the rollout), bit 1 = done too; a clear bit = device buffer + copy engine, slab by slab behind the
 * in-kernel progress flags.  Default 1 (fastest with one GPU per host; measured variants: csrc/env.cu, profiles/README.md).             */
int rb_env_set_host_transport(rb_env* env, int kernel_stores);
/* device pointers of the resident rollout buffer the last rb_env_rollout_policy_host call filled ([T,N,11], [T,N,4], [T,N], [T,N]);
 * any out pointer may be NULL.  Owned by the env, valid until the next host rollout or rb_env_destroy.  (reward / done are only
 * present on the device when the host buffers of that call were pageable or NULL.) */
int rb_env_rollout_buffer(rb_env* env, float** obs_buf_dev, float** pd_buf_dev, float** rew_buf_dev, uint8_t** done_buf_dev, int* T);

/* ------------------------------------------------------------------------------------------------ student -
 * Flat parameter layout: POLICY64 = the nout=4 policy layout above; MLP = for each layer W[in][out] then b[out],
 * layers 16-24-128-128-32-4.                                                                                 */
int64_t rb_student_param_count(int kind);
int rb_student_input_dim(int kind);
int64_t rb_student_workspace_bytes(int kind, int64_t batch, int device);

/* Student input assembly  src/distilation/mlp_train.py:50-52: x[B,16] = concat(dropout(ob, keep_prob), prev_pdflat, prev_rew)
 * dropout mask = floor(keep_prob + u), u Philox keyed (seed, sample_id0 + row, iteration); keep_prob >= 1: no dropout. */
int rb_student_mlp_input(const float* obs_dev, const float* prev_pdflat_dev, const float* prev_rew_dev, int64_t B,
                         float keep_prob, uint64_t seed, uint32_t sample_id0, uint32_t iteration, float* x_dev, void* stream);

/* sess.run(s_pdflat)  src/distilation/mlp_train.py:173-186 -- forward only.  x_dev[B,in] -> s_pdflat_dev[B,4] */
int rb_student_fwd(int kind, const float* params_dev, const float* x_dev, int64_t B, float* s_pdflat_dev, int mode, void* stream);
/* same with a workspace (rb_student_workspace_bytes() bytes) -- required for RB_MODE_TC, which folds the linear layer first */
int rb_student_fwd_ws(int kind, const float* params_dev, const float* x_dev, int64_t B, float* s_pdflat_dev, void* workspace_dev, int mode,
                      void* stream);

/* sess.run([loss, minimize_adam]) minus the Adam update  src/distilation/mlp_train.py:148-161 ;
 * lossandgrad  backup/student_rollout.py:646,708.
 * Forward + backward of the student with the loss (loss.py:3-13) and dL/dpdflat fused.
 *   gradloss_dev[P+1]: flat gradient (same layout as params) followed by the summed loss -> ONE all-reduce carries both.
 *   s_pdflat_dev[B,4] (may be NULL): student output (its mean half is the DAgger action).
 *   workspace_dev: rb_student_workspace_bytes() bytes.  Deterministic: fixed reduction order, no atomics.          */
int rb_student_loss_grad(int kind, const float* params_dev, const float* x_dev, const float* t_pdflat_dev, int64_t B,
                         int loss_kind, float* s_pdflat_dev, float* gradloss_dev, void* workspace_dev, int mode, void* stream);

/* adam.minimize  src/distilation/mlp_train.py:75-80 ; MpiAdam.update  backup/student_rollout.py:658,709
 * TF1 form: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; p -= lr_t*m/(sqrt(v)+eps).  grad is multiplied by grad_scale first
 * (1/world_size reproduces MpiAdam's averaged gradient).  step_t is 1-based.                                      */
int rb_adam_step(float* params_dev, float* m_dev, float* v_dev, const float* grad_dev, int64_t P, int64_t step_t, float lr,
                 float beta1, float beta2, float eps, float grad_scale, void* stream);

/* rb_student_loss_grad immediately followed by rb_adam_step on the same parameters (single rank: nothing to all-reduce).
 * RB_MODE_TC runs both inside ONE cooperative kernel launch.                                                     */
int rb_student_step(int kind, float* params_dev, float* m_dev, float* v_dev, const float* x_dev, const float* t_pdflat_dev, int64_t B,
                    int loss_kind, float* s_pdflat_dev, float* gradloss_dev, void* workspace_dev, int64_t step_t, float lr, float beta1,
                    float beta2, float eps, float grad_scale, int mode, void* stream);

/* Data-parallel form of rb_student_step (MpiAdam.update = Allreduce + Adam, backup/student_rollout.py:658,709), RB_MODE_TC only:
 * ONE cooperative kernel per rank computes the local [grad | loss], exchanges it with a one-shot all-reduce over NVLink peer
 * memory and applies Adam.  The exchange is a low-latency push: every rank stores {value, epoch} pairs (8 bytes, delivered atomically)
 * into every rank's receive area and sums its own area in rank order as the pairs arrive => bit-identical sums everywhere, no flag
 * round trip, no remote loads.
 *   peer_grad_slots[world]: device address, valid on THIS rank, of every rank's receive area for this step: uint64 [world][SL] with
 *                           SL = rb_student_param_count + 1 rounded up to 64 (symmetric / IPC-mapped memory, zero before the first
 *                           step; the caller alternates two areas per rank from step to step),
 *   peer_flags[world]     : reserved (may be NULL),
 *   epoch                 : 1, 2, 3, ... identical on all ranks.  gradloss_dev receives the all-reduced [grad | loss].
 * All ranks must launch the same step; the kernels wait for one another.                                               */
int rb_student_step_dp(int kind, float* params_dev, float* m_dev, float* v_dev, const float* x_dev, const float* t_pdflat_dev, int64_t B,
                       int loss_kind, float* s_pdflat_dev, float* gradloss_dev, void* workspace_dev, int64_t step_t, float lr, float beta1,
                       float beta2, float eps, float grad_scale, int rank, int world, const uint64_t* peer_grad_slots,
                       const uint64_t* peer_flags, uint32_t epoch, void* stream);

/* Debug aid: globaltimer stamps (ns) of CTA 0 at the phase boundaries of the last RB_MODE_TC student launch (48 values: [0,12) launch phases --
 * 0 start, 1 set up, 2 weight image in shared memory, 3 first tile, 4 tiles done, 5 partials written (+ fused env step), 6 after the first grid
 * barrier, 7 reduced, 8 after the second grid barrier, 9 un-fold [exchange] Adam done, 11 end; [16,31) phases of CTA 0's first tile; 34 gradient pushed
 * to the peers; see student_tc.cu).  Synchronises the device.                                                                          */
int rb_debug_student_timers(unsigned long long* host_out16);

/* ------------------------------------------------------------------------------------------------ DAgger --
 * One lock-step DAgger iteration pieces (src/distilation/mlp_train.py:143-204 batched):
 * rb_dagger_observe: for every env write ob[N,11], teacher label t_pdflat[N,4] and the student input x[N,in]
 *   (POLICY64: x = ob; MLP: x = [dropout(ob), prev teacher pdflat, prev recorded reward], dataset.py:118-143 semantics,
 *   zeros at the first step of an episode) and, when x_act_dev is not NULL (MLP student), x_act[N,16] = the same rows WITHOUT the
 *   observation dropout: the reference applies dropout to the TRAINING batch only (keep_prob = KEEP_PROB, mlp_train.py:151) and
 *   acts on the clean observation (keep_prob 1, mlp_train.py:171-186, lstm_train.py:176).
 * rb_dagger_act: step every env with the mean half of s_pdflat[N,4]; records rew[N], done[N]; keeps prev-pdflat / prev-rew. */
typedef struct rb_dagger rb_dagger;
int rb_dagger_create(rb_dagger** out, rb_env* env, int student_kind, float keep_prob);
int rb_dagger_destroy(rb_dagger* d);
int rb_dagger_observe(rb_dagger* d, const float* teacher_params_dev, uint32_t iteration, float* obs_dev, float* t_pdflat_dev,
                      float* x_dev, float* x_act_dev, int mode, void* stream);
/* Checkpoint / resume of the loop (`train(train, restore)`: main.py:24-27, lstm_train.py:86-87,102-107 save / restore the student every episode):
 * the per-env state the handle carries between iterations -- teacher pdflat of the previous record [N,4], `rew` field of the previous record [N],
 * reward of the last env.step [N] (dataset.py:118-143 `prev` / `prew`).  Device pointers, device-to-device copies on `stream`; NULL skips a field.
 * Together with rb_env_get_state / rb_env_set_state, the student parameters + Adam moments and the iteration counter this resumes a run
 * bit-exactly.                                                                                                                          */
int rb_dagger_get_state(rb_dagger* d, float* prev_t_pdflat_dev, float* prev_rec_rew_dev, float* last_reward_dev, void* stream);
int rb_dagger_set_state(rb_dagger* d, const float* prev_t_pdflat_dev, const float* prev_rec_rew_dev, const float* last_reward_dev, void* stream);
/* RB_MODE_TC observe caches the split-weight image of the (frozen, teacher.py:17-20) teacher keyed by the parameter POINTER;
 * call this after modifying the teacher parameters in place.                                                                */
int rb_dagger_invalidate_teacher(rb_dagger* d);
/* One whole DAgger iteration (RB_MODE_TC): rb_dagger_observe + rb_student_step[_dp] + rb_dagger_act, with every per-step quantity
 * (dropout iteration, Adam step, exchange epoch / slot parity) read from a device-side clock set once by rb_dagger_set_clock and
 * advanced at the end of the iteration -- so the launches are captured ONCE in a CUDA graph (use_graph != 0) and replayed with a single
 * cudaGraphLaunch per iteration.  The env step (rb_dagger_act's work) runs inside the student launch (CTAs that would idle at its first
 * grid barrier step the envs whose forward pass is published), so an iteration is two launches.  x_act_dev (MLP student, may be NULL):
 * un-dropped input rows; when given, every tile of the student launch runs the forward pass twice -- on x_act for s_pdflat_dev, the
 * pdflat the envs are stepped with, then on x for the loss and the gradient (NULL: one pass, the student acts on its training input).
 * world > 1: slots_even / slots_odd / flags as in rb_student_step_dp (the two receive-area sets alternate). */
int rb_dagger_set_clock(rb_dagger* d, uint32_t iteration, uint32_t adam_step, uint32_t epoch, void* stream);
int rb_dagger_step(rb_dagger* d, const float* teacher_params_dev, float* params_dev, float* m_dev, float* v_dev, float* gradloss_dev,
                   void* workspace_dev, float* obs_dev, float* t_pdflat_dev, float* x_dev, float* x_act_dev, float* s_pdflat_dev, float* rew_dev,
                   uint8_t* done_dev,
                   int loss_kind, float lr, float beta1, float beta2, float eps, float grad_scale, int rank, int world,
                   const uint64_t* slots_even, const uint64_t* slots_odd, const uint64_t* flags, int use_graph, void* stream);
int rb_dagger_act(rb_dagger* d, const float* s_pdflat_dev, const float* t_pdflat_dev, float* rew_dev, uint8_t* done_dev,
                  void* stream);
/* The per-iteration loss the reference prints (src/distilation/mlp_train.py:148-161, 199-201) without a stream synchronise: the last
 * kernel of rb_dagger_step stores {loss, iterations done} as one 8-byte word into page-locked mapped host memory; this call polls it
 * until the device clock has reached `iteration` (= the value given to rb_dagger_set_clock + the number of rb_dagger_step calls since)
 * and returns that iteration's loss (summed over ranks when world > 1).  Fails if a later iteration has already overwritten it. */
int rb_dagger_wait_loss(rb_dagger* d, uint32_t iteration, float* loss_host_out);

/* ------------------------------------------------------------------------------------------------ LSTM student -
 * student_lstm_graph  src/distilation/student_nn.py:21-49 (built at lstm_train.py:32-57): dropout(ob) (+) dense32(prev_pdflat) ->
 * LSTMCell(200) -> per-unrolled-step heads 64-128-64-32-4, T = STEPS_UNROLLED = 10.  Window tensors are time-major [T,B,.].
 * state = [2,B,200] (c, m) -- NULL means zeros (training windows, lstm_train.py:159); acting carries it (lstm_train.py:171-182).
 * Flat parameters: rb_lstm_param_count() floats, layout W_e[4][32] b_e[32] W_l[243][800] b_l[800] then per step tau:
 * W1[200][64] b1 W2[64][128] b2 W3[128][64] b3 W4[64][32] b4 W5[32][4] b5.   All GEMMs run on tcgen05 (bf16x3).              */
int64_t rb_lstm_param_count(void);
int rb_lstm_steps(void);
int rb_lstm_units(void);
int64_t rb_lstm_workspace_bytes(int64_t batch);
/* sess.run((s_pdflat, final_state_batch))  lstm_train.py:171-182: forward only, keep_prob = 1 */
int rb_lstm_fwd(const float* params_dev, const float* ob_dev, const float* prev_pdflat_dev, const float* init_state_dev, int64_t B,
                float* s_pdflat_dev, float* final_state_dev, void* workspace_dev, void* stream);
/* sess.run([loss, minimize_adam]) minus Adam  lstm_train.py:145-160: forward, KL loss (loss.py:3-13), back-propagation through time.
 * gradloss_dev[P+1] = flat gradient | summed loss.  Dropout mask: Philox keyed (seed, sample_id0 + row, iteration).             */
int rb_lstm_loss_grad(const float* params_dev, const float* ob_dev, const float* prev_pdflat_dev, const float* t_pdflat_dev,
                      const float* init_state_dev, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, uint32_t iteration,
                      int loss_kind, float* s_pdflat_dev, float* final_state_dev, float* gradloss_dev, void* workspace_dev, void* stream);
/* rb_lstm_loss_grad + TF-form Adam as ONE CUDA-graph launch (~250 kernels captured once; the dropout iteration and the Adam step come
 * from a device-side clock owned by the context, advanced by the last kernel of the graph).                                     */
typedef struct rb_lstm_ctx rb_lstm_ctx;
int rb_lstm_ctx_create(rb_lstm_ctx** out, int device);
int rb_lstm_ctx_destroy(rb_lstm_ctx* ctx);
int rb_lstm_ctx_set_clock(rb_lstm_ctx* ctx, uint32_t iteration, uint32_t adam_step, void* stream);
int rb_lstm_step(rb_lstm_ctx* ctx, float* params_dev, float* m_dev, float* v_dev, const float* ob_dev, const float* prev_pdflat_dev,
                 const float* t_pdflat_dev, const float* init_state_dev, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0,
                 int loss_kind, float* s_pdflat_dev, float* gradloss_dev, void* workspace_dev, float lr, float beta1, float beta2, float eps,
                 float grad_scale, int use_graph, void* stream);
/* ------------------------------------------------------------------------------------------------ two-headed LSTM student -
 * lstm_graph + total loss of the backup experiment  src/distilation/backup/student_rollout.py:130-200 (graph), :303-328 (loss = KL(student ||
 * teacher) + sum (reward - reward_target)^2), :331-338 (Adam 1e-3): a shared LSTMCell(units) over `steps` unrolled steps of
 * [dropout(ob) (11) | stepped action (2)], and per unrolled step an un-shared head: trunk = tanh(dense_trunk(m)); reward = dense_1 over
 * n_reward_hidden tanh layers on the trunk; pdflat = dense_4(tanh(dense_action_hidden(trunk))).
 * spec = RB_LSTM2_SPEC_LEN ints {units, steps, carry_state, trunk, action_hidden, n_reward_hidden, reward_hidden[0..3]}:
 *   the checked-in source is {NUM_UNITS, STEPS_UNROLLED, 0, 128, 64, 1, 64} -- its loop never reassigns `state` (:156), so every unrolled step
 *   starts from the fed initial state and final_state is that state; the graph in the reference's tfevents files (src/~/reacher/data/viz/1) is
 *   {1, 2, 1, 128, 64, 3, 64, 32, 64} with the state carried through the unroll.
 * Windows are time-major: ob [T,B,11], action [T,B,2], t_pdflat / s_pdflat [T,B,4], reward / reward_target [T,B]; state [2,B,units] (c, m),
 * NULL = zeros.  Flat parameters: W_l[13+units][4 units] b_l, then per step: Wd bd | Wr_k br_k ... | Wro bro | Wa ba | Wp bp (row-major
 * [in][out] kernels, as tf.layers.dense).  gradloss = P + 3 floats: flat gradient | total loss | KL part | reward part.
 * Adam: rb_adam_step on the flat vectors.  All products run on tcgen05 (rb_gemm_bf16x3); results are bit-reproducible.               */
#define RB_LSTM2_SPEC_LEN 10
int64_t rb_lstm2_param_count(const int* spec);
int64_t rb_lstm2_workspace_bytes(const int* spec, int64_t batch);
/* sess.run((s_ac, final_state_combined))  backup/student_rollout.py:527-535: forward only, keep_prob = 1 */
int rb_lstm2_fwd(const int* spec, const float* params_dev, const float* ob_dev, const float* action_dev, const float* init_state_dev, int64_t B,
                 float* s_pdflat_dev, float* reward_dev, float* final_state_dev /* may be NULL */, void* workspace_dev, void* stream);
/* sess.run([loss, minimize_adam]) minus Adam  backup/student_rollout.py:508-522 */
int rb_lstm2_loss_grad(const int* spec, const float* params_dev, const float* ob_dev, const float* action_dev, const float* t_pdflat_dev,
                       const float* reward_target_dev, const float* init_state_dev, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0,
                       uint32_t iteration, int loss_kind, float* s_pdflat_dev, float* reward_dev, float* final_state_dev /* may be NULL */,
                       float* gradloss_dev, void* workspace_dev, void* stream);
/* rb_lstm2_loss_grad + TF-form Adam as ONE CUDA-graph launch (context and device-side clock of rb_lstm_step: rb_lstm_ctx_create /
 * rb_lstm_ctx_set_clock); pass the same (static) tensors every step, the graph is re-captured when a pointer or scalar changes.      */
int rb_lstm2_step(rb_lstm_ctx* ctx, const int* spec, float* params_dev, float* m_dev, float* v_dev, const float* ob_dev, const float* action_dev,
                  const float* t_pdflat_dev, const float* reward_target_dev, const float* init_state_dev, int64_t B, float keep_prob, uint64_t seed,
                  uint32_t sample_id0, int loss_kind, float* s_pdflat_dev, float* reward_dev, float* gradloss_dev, void* workspace_dev, float lr,
                  float beta1, float beta2, float eps, float grad_scale, int use_graph, void* stream);
/* The tensor-core GEMM the LSTM is built from: C[M,N] (+)= epilogue(A[M,K] B[K,N]), fp32 in/out, bf16x3 inside.
 * x_mn = 0: element (row, k) of the operand at X[row * ld + k]; 1: at X[k * ld + row].  epilogue: + bias[n], tanh (act = 1),
 * * (1 - H[m,n]^2).  workspace (optional, floats) enables deterministic split-K.                                              */
int rb_gemm_bf16x3(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, float* C, int ldc, int M, int N, int K,
                   const float* bias, int act, int accumulate, const float* H, int ldh, float* workspace, int64_t workspace_floats,
                   void* stream);
/* Pipeline shape of rb_gemm_bf16x3 (results are bit-identical either way): 0 = deep operand pipeline, one (128-wide tiles) or two CTAs per
 * SM; 1 = packed, a two-stage pipeline at two / three CTAs per SM; -1 (default) = packed whenever the grid holds more CTAs than the GPU has
 * SMs.  Process-wide tuning knob for measurements (scripts/r02/gemm_packing_ab.py); not part of any reference call path.              */
int rb_gemm_set_cta_packing(int mode);

/* ------------------------------------------------------------------------------------------------ dense stacks -
 * The auxiliary objectives of the reference's backup experiments, as a generic dense stack on rb_gemm_bf16x3:
 *   value-function regressor   src/distilation/backup/student_rollout_mlp_vf.py:251-276 ([prev_ob | next_ac] -> 64 linear -> 10 x tanh(100) -> 1,
 *                              v_loss = sum (vpred - vtarg)^2 :276, Adam lr 1e-2 :290-295), targets by add_vtarg :608-616
 *   reward-prediction head     src/distilation/backup/student_rollout.py:161-164 (64 tanh -> 1), loss += sum (reward - target)^2 :328
 *   KL-trained dense students of other widths (loss.py:3-13, backup/student_rollout.py:639-642)
 * dims[0..n_layers] are the widths (input first); acts[l] = 0 linear / 1 tanh for layer l (NULL: tanh everywhere but the last layer; the
 * output layer must be linear for rb_dense_loss_grad).  Flat parameters: for every layer W[d_in][d_out] (row-major) then b[d_out].
 * x [B, dims[0]], target / out [B, dims[n_layers]] row-major.  gradloss = rb_dense_param_count() + 1 floats: flat gradient | loss.
 * Results are bit-reproducible (fixed-order reductions).  Adam: rb_adam_step on the flat vectors.                                   */
#define RB_DENSE_MAX_LAYERS 16
int64_t rb_dense_param_count(int n_layers, const int* dims);
int64_t rb_dense_workspace_bytes(int n_layers, const int* dims, int64_t batch);
int rb_dense_fwd(const float* params_dev, int n_layers, const int* dims, const int* acts, const float* x_dev, int64_t B, float* out_dev,
                 void* workspace_dev, void* stream);
int rb_dense_loss_grad(const float* params_dev, int n_layers, const int* dims, const int* acts, const float* x_dev, const float* target_dev, int64_t B,
                       int loss_kind, float* out_dev /* may be NULL */, float* gradloss_dev, void* workspace_dev, void* stream);
/* add_vtarg  backup/student_rollout_mlp_vf.py:608-616 for `episodes` reward rows [episodes, steps]: vtarg[e][steps-1] = gamma^steps r[steps-1],
 * vtarg[e][i] = gamma^i r[i] + vtarg[e][i+1] (the reference's absolute-time discount, exponent quirk of the last step included).         */
int rb_vf_targets(const float* rew_dev, int64_t episodes, int steps, float gamma, float* vtarg_dev, void* stream);

/* ------------------------------------------------------------------------------------------------ dataset -
 * Device-resident rollout buffer with the semantics of the reference Dataset (src/distilation/dataset.py:72-296) for N lock-step
 * envs: a ring of `generations` x 50 steps x N records {ob[11], rew, t[4], s[4], with}; prev / prew of record j are t / rew of
 * record j-1 (zeros at j = 0: pdflat_at / rew_at, dataset.py:151-164).  The cursor (step in episode, generation) lives on the host;
 * all data movement is asynchronous on `stream`.                                                                          */
typedef struct rb_dataset rb_dataset;
int rb_dataset_create(rb_dataset** out, int64_t num_envs, int64_t generations, int device);
int rb_dataset_destroy(rb_dataset* d);
/* Dataset.write  dataset.py:118-143: append one record per env.  rew / t / s may be NULL (zeros).  stepped_with: 0 = 't', 1 = 's'. */
int rb_dataset_write(rb_dataset* d, const float* ob_dev, const float* rew_dev, const float* t_pdflat_dev, const float* s_pdflat_dev,
                     int stepped_with, void* stream);
/* Dataset.flush  dataset.py:146-149: close the N current episodes (they must hold exactly 50 records, as in the reference). */
int rb_dataset_flush(rb_dataset* d);
int64_t rb_dataset_num_episodes(const rb_dataset* d);   /* Dataset.num_episodes: episodes flushed so far                       */
int64_t rb_dataset_num_available(const rb_dataset* d);  /* complete episodes still in the ring                                 */
int rb_dataset_episode_len(const rb_dataset* d);        /* len(curr_episode)                                                   */
/* Copy one flushed generation (absolute index, must still be in the ring) to HOST buffers ob[50,N,11] rew[50,N] t[50,N,4] s[50,N,4]
 * with[50,N] -- what Dataset.dump (dataset.py:80-85 -> DatasetStore.store :31-40) serialises into gzip-JSON pages.  Synchronises.  */
int64_t rb_dataset_generations(const rb_dataset* d);
int rb_dataset_export_host(rb_dataset* d, int64_t generation, float* ob_host, float* rew_host, float* t_host, float* s_host, uint8_t* with_host);
/* Whole-ring snapshot / restore (exact resume of the loops that train from the Dataset; the reference restores its student every episode,
 * lstm_train.py:86-87,102-107 -- here the replay buffer comes back too): rb_dataset_ring_rows() rows of every field (ob [rows,11], rew [rows],
 * t / s [rows,4], with [rows]) + the cursor (records in the open episodes, generations flushed).  HOST buffers; synchronises.       */
int64_t rb_dataset_ring_rows(const rb_dataset* d);
int rb_dataset_save_host(rb_dataset* d, float* ob_host, float* rew_host, float* t_host, float* s_host, uint8_t* with_host, int* step_out,
                         int64_t* generations_out);
int rb_dataset_load_host(rb_dataset* d, const float* ob_host, const float* rew_host, const float* t_host, const float* s_host,
                         const uint8_t* with_host, int step, int64_t generations);
/* Dataset.training_batches  dataset.py:179-210: B episodes drawn with replacement and ONE shared start in [0, 50-T], Philox keyed
 * (seed; draw, b).  Time-major outputs ob[T,B,11], t[T,B,4], prev[T,B,4], prew[T,B,1]; episodes_out[B] / start_out[1] (optional)
 * report what was drawn.                                                                                                      */
int rb_dataset_training_batch(rb_dataset* d, uint64_t seed, uint32_t draw, int B, int T, float* ob_out_dev, float* t_out_dev, float* prev_out_dev,
                              float* prew_out_dev, int32_t* episodes_out_dev, int32_t* start_out_dev, void* stream);
/* Dataset.test_batch  dataset.py:213-290, for every env at once: ob[T,N,11] = last T-1 observations of the current episode
 * (left zero padded) + ob_cur; prev[T,N,4] / prew[T,N,1] = t / rew of the records one step earlier (zero padded).  The reference
 * places the single env's window in the LAST batch row of a zero batch; env.py does that for num_envs == 1.                  */
int rb_dataset_test_batch(rb_dataset* d, const float* ob_cur_dev, int T, float* ob_out_dev, float* prev_out_dev, float* prew_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REACHER_B200_H */
