#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode tc|fp32]
  N > 1: launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Headline workload = BASELINE.json config 3 per GPU: 65 536 lock-step Reacher-v2 envs, fused teacher MLP (2x64 tanh) in the
loop, device-resident rollout buffer; one bench "step" = one 50-step chunk (one episode of every env) = 3 276 800 env-steps
per GPU.  Weak scaling: every rank owns 65 536 envs (global env ids rank*65536 ..), no data-path collective.
`value`   env-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     same metric through the host-buffer C-ABI call (rb_env_rollout_policy_host): H2D of the teacher parameters and D2H of
          the step's result (reward + done of every env-step) inside the timed region; the rollout buffer (obs, pdflat) is written
          and stays on the device.  `e2e_full_buffer`: the same call bringing the whole buffer to the host (PCIe-bound).
`distill` BASELINE.json config 4 shard (32 768 envs per GPU): DAgger iterations = env step + teacher label + student
          forward/backward + KL + [NCCL all-reduce of the flat gradient] + Adam; samples/s == env-steps/s of that loop.
`step_api` the gym-style single-step kernel (HBM-bound) at 4 194 304 envs.
`config1`  BASELINE.json config 1 (one env, 1000-step teacher rollout): host-thread restatement vs the gym surface at batch 1 vs one fused launch.
`cpu_baseline` / --impl reference: the float64 C restatement of the reference's CPU path (oracle/, OpenMP over host cores) on a
          bounded sample of the same workload.  The reference itself (TF-1.10 + gym + MuJoCo-1.50) is not installable here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
CHUNK_T = 50
DISTILL_ENVS_PER_GPU = 32768
STEP_API_ENVS = 1 << 22
ALG_BYTES_ROLLOUT = 65.0      # fused rollout writes one buffer row per env-step: ob 44 + pdflat 16 + rew 4 + done 1 B (17 scalars)
ROLLOUT_TRAFFIC_NCU = 160.0e6   # dram__bytes_read.sum + dram__bytes_write.sum of one k_rollout_policy_tc launch (65 536 envs x 50 steps),
                                # profiles/r01_ncu_full_k_rollout_policy_tc_final.csv; the rest of the 213 MB is still dirty in L2 at kernel end
MUFU_PER_ENV_STEP = 169.0       # 128 tanh x 1.25 (ex2 each, one rcp per four) + 8 rcp + 1 sqrt
XU_LANES_PER_CLK_PER_SM = 16.0  # B200 MUFU rate
ALG_BYTES_STEP = 113.0        # single-step API, I/O 57 B + state round trip 56 B
STEP_TRAFFIC_NCU = 515.4e6    # dram read + write of one k_step launch at 4 194 304 envs (profiles/r01_ncu_full_k_step.csv) = 122.9 B per env-step
STUDENT_TRAFFIC_NCU = 4.46e6  # dram read + write of one k_student_tc<SpecMLP> launch of the DAgger graph at 32 768 samples
                              # (profiles/r01_ncu_full_k_student_tc_final.csv): inputs and env state read once, everything else stays in L2
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
FLOP_PER_ENV_STEP = 450.0 + 9856.0          # physics + teacher MLP
FLOP_PER_SAMPLE = {"mlp": 144.4e3, "policy64": 30.3e3}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        if not [1 for (t, _) in self.lines if t0 <= t <= t1]:
            time.sleep(0.2)                               # very short timed region: take the samples closest to it
        rows = [l for (t, l) in self.lines if t0 <= t <= t1]
        if not rows:
            rows = [l for (_, l) in sorted(self.lines, key=lambda tl: abs(tl[0] - 0.5 * (t0 + t1)))[:3]]
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))

    def stop(self):
        if self.proc:
            self.proc.kill()


def cpu_reference_leg(seconds=12.0, nthreads=0):
    """Time the C restatement of the reference's CPU rollout (teacher in the loop) on the host cores, bounded sample."""
    import numpy as np
    from oracle import reacher_c as RC
    from reacherdistilation_b200.teacher import init_policy_params
    p = init_policy_params(seed=0)
    cores = RC.max_threads() if nthreads <= 0 else nthreads
    cal = RC.ReacherOracleC(2048, seed=0, nthreads=nthreads); cal.reset()
    t0 = time.perf_counter(); cal.rollout_policy(CHUNK_T, p, record=False); dt = time.perf_counter() - t0
    rate = 2048 * CHUNK_T / dt
    n = int(min(ENVS_PER_GPU, max(2048, rate * seconds / CHUNK_T)))
    env = RC.ReacherOracleC(n, seed=0, nthreads=nthreads); env.reset()
    obs, pd, rw, dn = (np.zeros((CHUNK_T, n, 11)), np.zeros((CHUNK_T, n, 4)), np.zeros((CHUNK_T, n)), np.zeros((CHUNK_T, n), np.uint8))
    reps, t0 = 0, time.perf_counter()
    while reps == 0 or time.perf_counter() - t0 < seconds:          # bounded sample: ~`seconds` of CPU work on all host cores
        env.rollout_policy(CHUNK_T, p, record=True)
        reps += 1
    dt = time.perf_counter() - t0
    return dict(value=n * CHUNK_T * reps / dt, unit="env-steps/s", cores=cores, kind="port",
                sample="%d x (%d envs x %d steps) teacher-in-the-loop rollout chunks, float64 C restatement (oracle/reacher_oracle.c, OpenMP), %.1f s"
                       % (reps, n, CHUNK_T, dt)), n, dt


def cpu_distill_leg(kind="mlp", seconds=6.0):
    """numpy float64 restatement of one DAgger optimiser step (student fwd/bwd + KL + Adam) on a bounded batch."""
    import numpy as np
    from oracle import nn_np as NN
    rng = np.random.default_rng(0)
    B = 8192
    if kind == "mlp":
        P = (rng.standard_normal(NN.mlp_param_count()) * 0.1).astype(np.float32)
        x = rng.standard_normal((B, 16))
    else:
        P = (rng.standard_normal(NN.policy_param_count(4)) * 0.1).astype(np.float32); P[:11] = 0; P[11:22] = 1
        x = rng.standard_normal((B, 11))
    t = rng.standard_normal((B, 4)) * 0.3
    opt = NN.AdamTF(P.size)
    theta = P.astype(np.float64)
    n_it, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        if kind == "mlp":
            s, hs = NN.mlp_fwd(x, theta.astype(np.float32)); l, ds = NN.kl_loss(s, t); g = NN.mlp_bwd(hs, theta.astype(np.float32), ds)
        else:
            s = NN.policy_fwd(x, theta.astype(np.float32), nout=4); l, ds = NN.kl_loss(s, t); g = NN.policy_bwd(x, theta.astype(np.float32), ds)
        theta = opt.update(theta, g)
        n_it += 1
    dt = time.perf_counter() - t0
    return dict(value=B * n_it / dt, unit="samples/s", cores=os.cpu_count(), kind="port",
                sample="%d optimiser steps of batch %d, numpy float64 restatement (oracle/nn_np.py), %.1f s" % (n_it, B, dt))


def config1_leg(steps=1000):
    """BASELINE config 1, the reference's own CPU-runnable case (SURVEY 8(d).1 / CPU plan (i)): ONE env, a 1000-step teacher rollout in the
    per-step loop shape of mlp_train.py:120-139 -- (cpu) the restatement on one host thread, (gpu_gym_loop) the drop-in gym surface at batch 1
    (one rb_policy_fwd + one rb_env_step_host per step: launch-latency bound, reported for honesty), (gpu_fused) the same 1000 steps as ONE
    rb_env_rollout_policy launch with the result read back.  A context figure; batch 1 is not what the GPU path is built for."""
    import numpy as np
    import torch
    from oracle import nn_np as NN
    from oracle import reacher_c as RC
    from reacherdistilation_b200 import MODE_FP32
    from reacherdistilation_b200.env import VecReacher, make_mujoco_env
    from reacherdistilation_b200.teacher import TeacherAgent, init_policy_params
    p = init_policy_params(seed=0)
    c = RC.ReacherOracleC(1, seed=0, nthreads=1)
    ob = c.reset()
    t0 = time.perf_counter()
    for _ in range(steps):
        flat = NN.policy_fwd(ob.astype(np.float32), p)
        ob, r, d = c.step(flat[:, :2])
    cpu = steps / (time.perf_counter() - t0)
    env = make_mujoco_env("Reacher-v2", 0)
    teacher = TeacherAgent(env, params=p, mode=MODE_FP32)
    ob = env.reset()
    for rep in range(2):                                   # first pass = warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(steps):
            mean, _flat = teacher.mean_and_flat(ob)
            ob, r, new, _ = env.step(mean.cpu().numpy())
        gym_loop = steps / (time.perf_counter() - t0)
    env.close()
    v = VecReacher(num_envs=1, seed=0); v.reset()
    pd = torch.from_numpy(p).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = v.rollout_policy(pd, steps, mode=MODE_FP32)
        ret = float(out["rew"].sum())                      # device -> host read of the result
        fused = steps / (time.perf_counter() - t0)
    v.close()
    return dict(workload="config1: 1 env, %d-step teacher rollout (20 episodes)" % steps, unit="env-steps/s", cpu_restatement_1_thread=cpu,
                gpu_gym_loop_batch1=gym_loop, gpu_fused_one_launch=fused, teacher_return_per_episode=ret / (steps / 50.0))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restated; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    import numpy as np
    from oracle import reacher_c as RC
    from reacherdistilation_b200.teacher import init_policy_params
    p = init_policy_params(seed=0)
    cores = RC.max_threads()
    cal = RC.ReacherOracleC(2048, seed=0); cal.reset()
    t0 = time.perf_counter(); cal.rollout_policy(CHUNK_T, p, record=False); rate = 2048 * CHUNK_T / (time.perf_counter() - t0)
    budget = 120.0 / (steps + warm)                                  # whole run within a few minutes
    n = int(min(ENVS_PER_GPU, max(1024, rate * min(budget, 10.0) / CHUNK_T)))
    env = RC.ReacherOracleC(n, seed=0); env.reset()
    for _ in range(warm):
        env.rollout_policy(CHUNK_T, p, record=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        env.rollout_policy(CHUNK_T, p, record=True)
    dt = time.perf_counter() - t0
    val = n * CHUNK_T * steps / dt
    sample = "%d of %d envs per step x %d steps/chunk, float64 C restatement of the MuJoCo+TF CPU path, OpenMP %d threads" % (n, ENVS_PER_GPU, CHUNK_T, cores)
    line = dict(impl="reference", metric="reacher_env_steps_per_sec", value=val, unit="env-steps/s", n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="config3: %d envs/GPU, fused teacher MLP (11-64-64-2 tanh) in the loop, 50-step chunk per step, "
                                     "rollout buffer device-resident" % ENVS_PER_GPU, envs_per_gpu=ENVS_PER_GPU, chunk_steps=CHUNK_T,
                            reference_sample="each step = %d of the %d envs x %d steps on the host cores (bounded sample)" % (n, ENVS_PER_GPU, CHUNK_T)),
                cpu_baseline=dict(value=val, unit="env-steps/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fp32", "tc"])
    ap.add_argument("--student", default="mlp", choices=["mlp", "policy64"])
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements (distill, step API, CPU legs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from reacherdistilation_b200 import MODE_FP32, MODE_TC, STUDENT_MLP, STUDENT_POLICY64, _lib
    from reacherdistilation_b200.dist import init_from_env, max_over_ranks
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    from reacherdistilation_b200.teacher import init_policy_params

    rank, world, local = init_from_env()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, W = max(1, args.steps), max(3, args.warmup)
    L = _lib.lib()
    mode = {"fp32": MODE_FP32, "tc": MODE_TC}.get(args.mode) if args.mode != "auto" else (MODE_TC if L.rb_mode_available(MODE_TC) else MODE_FP32)
    mode_name = "tc(tcgen05 bf16x3)" if mode == MODE_TC else "fp32(cuda cores)"
    pk = peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        """barrier+sync, CUDA events on the launching stream, barrier+sync; returns max-over-ranks seconds and host window."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.time()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        h1 = time.time()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3, dev), (h0, h1)

    sampler = ClockSampler(local) if rank == 0 else None
    n = ENVS_PER_GPU
    teacher = torch.from_numpy(init_policy_params(seed=0)).to(dev)
    env = VecReacher(num_envs=n, seed=0, device=local, env_offset=rank * n)
    env.reset()
    buf = dict(obs=torch.empty((CHUNK_T, n, 11), device=dev), pdflat=torch.empty((CHUNK_T, n, 4), device=dev),
               rew=torch.empty((CHUNK_T, n), device=dev), done=torch.empty((CHUNK_T, n), dtype=torch.uint8, device=dev))
    step_fn = lambda: env.rollout_policy(teacher, CHUNK_T, nout=2, mode=mode, out=buf)
    for _ in range(W):
        step_fn()
    sec, win = timed(step_fn, K)
    env_steps = float(n) * CHUNK_T * K * world
    value = env_steps / sec
    kernel_s = sec / K                                                 # one kernel launch per step
    clocks = sampler.window(*win) if sampler else None
    mean_rew = float(buf["rew"].mean())

    # ---- e2e: the host-buffer C-ABI call (rb_env_rollout_policy_host), synchronous, copies inside the timed region -------------
    #   e2e             : H2D of the teacher parameters, the rollout (buffer stays device-resident for the distillation loop, north_star (3)),
    #                     D2H of the step's RESULT: reward and done of every env-step (what the reference prints / extract_reward.py consumes)
    #   e2e_full_buffer : same call returning the WHOLE rollout buffer (obs, pdflat, reward, done = 65 B/env-step) to the host -- PCIe-bound
    Ke = max(3, min(K, 20))
    hbuf = dict(obs=torch.empty((CHUNK_T, n, 11)).pin_memory(), pdflat=torch.empty((CHUNK_T, n, 4)).pin_memory(),
                rew=torch.empty((CHUNK_T, n)).pin_memory(), done=torch.empty((CHUNK_T, n), dtype=torch.uint8).pin_memory())
    hres = dict(obs=None, pdflat=None, rew=hbuf["rew"], done=hbuf["done"])
    tparams_host = torch.from_numpy(init_policy_params(seed=0)).pin_memory()

    def e2e_time(out):
        fn = lambda: env.rollout_policy_host(tparams_host, CHUNK_T, nout=2, mode=mode, out=out)
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            fn()                                                       # synchronous call (copies + sync inside)
        barrier()
        return max_over_ranks(time.perf_counter() - t0, dev)
    e2e_sec, e2e_full_sec = e2e_time(hres), e2e_time(hbuf)
    e2e = dict(value=float(n) * CHUNK_T * Ke * world / e2e_sec, unit="env-steps/s", h2d_bytes_per_step=int(tparams_host.numel() * 4),
               d2h_bytes_per_step=int(n * CHUNK_T * (4 + 1)), steps=Ke, api="rb_env_rollout_policy_host (VecReacher.rollout_policy_host)",
               result="reward[T,N] f32 + done[T,N] u8 to pinned host memory (reward written by the kernel into the mapped buffer, done "
                      "copied after it); obs / pdflat are written to the device-resident rollout buffer (rb_env_rollout_buffer) and stay there",
               mean_reward_host=float(hbuf["rew"].mean()))
    e2e_full = dict(value=float(n) * CHUNK_T * Ke * world / e2e_full_sec, unit="env-steps/s", h2d_bytes_per_step=int(tparams_host.numel() * 4),
                    d2h_bytes_per_step=int(n * CHUNK_T * (44 + 16 + 4 + 1)), steps=Ke, note="whole rollout buffer to the host: PCIe-bound")
    env.close()
    del buf, hbuf

    line = dict(metric="reacher_env_steps_per_sec", value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=W, ms_per_step=1e3 * sec / K,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="config3: %d envs/GPU, fused teacher MLP (11-64-64-2 tanh) in the loop, 50-step chunk per step, "
                                     "rollout buffer device-resident" % n, envs_per_gpu=n, chunk_steps=CHUNK_T, policy_mode=mode_name,
                            l2="no flush: each step writes a fresh %.0f MB rollout buffer (> 126 MB L2); state 2.6 MB stays in registers"
                               % (n * CHUNK_T * 68 / 1e6), seed=0, mean_teacher_reward=mean_rew),
                e2e=e2e, e2e_full_buffer=e2e_full, gpu_launches=K, clocks=clocks)
    line["roofline"] = dict(bound="hbm", achieved=ALG_BYTES_ROLLOUT * n * CHUNK_T / kernel_s / 1e9, peak=pk["hbm"], unit="GB/s",
                            frac=ALG_BYTES_ROLLOUT * n * CHUNK_T / kernel_s / 1e9 / pk["hbm"],
                            traffic=(ROLLOUT_TRAFFIC_NCU if (mode == MODE_TC and n == 65536) else None), peak_source=pk["src"],
                            kernel="k_rollout_policy_%s" % ("tc" if mode == MODE_TC else "fp32"),
                            note="HBM is NOT the limiter of the fused rollout (it only writes the 65 B buffer row per env-step, state stays in "
                                 "registers, SURVEY 8(d)); the binding resources are instruction issue and the XU (MUFU) pipe: see `pipes`")
    sms = L.rb_sm_count(local)
    xu_ceiling = sms * XU_LANES_PER_CLK_PER_SM * 1.965e9 / MUFU_PER_ENV_STEP
    line["pipes"] = dict(xu_mufu_per_env_step=MUFU_PER_ENV_STEP, xu_ceiling_env_steps_per_s=xu_ceiling, frac_of_xu_ceiling=value / world / xu_ceiling,
                         tensor_tflops_bf16x3=3 * 9856.0 * n * CHUNK_T / kernel_s / 1e12, tensor_peak_tflops=pk["bf16_burst"],
                         physics_flop_per_env_step=450.0, teacher_flop_per_env_step=9856.0,
                         note="XU ceiling = SMs x 16 MUFU lanes/clk x 1.965 GHz / 169 MUFU per env-step; ncu (profiles/README.md): XU pipe 40 % of active "
                              "cycles, issue slots 57 %, tensor pipe 18 %; only 13.8 warps/SM exist at 65 536 envs (latency-bound, see DESIGN.md 4.1)")

    if not args.quick:
        # ---- distill: DAgger iterations on the config-4 shard ---------------------------------------------------
        kind = STUDENT_MLP if args.student == "mlp" else STUDENT_POLICY64
        nd = DISTILL_ENVS_PER_GPU
        tr = DaggerTrainer(num_envs=nd, seed=0, device=local, student_kind=kind, mode=mode, env_offset=rank * nd)
        tr.sync_params()
        Kd = max(20, min(K, 200))
        for _ in range(W):
            tr.step()
        dsec, _ = timed(tr.step, Kd)
        # student kernel alone (device time of the dominant kernel of this loop)
        lg = lambda: tr.student.loss_grad(tr.x, tr.t_pd, s_out=tr.s_pd)      # ONE cooperative launch in tc mode (fold, tiles, reduce, un-fold)
        for _ in range(3):
            lg()
        ksec, _ = timed(lg, 20)
        # e2e: the loss comes back to the host every iteration
        def dstep_e2e():
            tr.step()
            tr.wait_loss()        # graph path: the iteration's last kernel posts {loss, iteration} into mapped host memory, the host polls it
        barrier(); t0 = time.perf_counter()
        for _ in range(Kd):
            dstep_e2e()
        barrier()
        de2e = max_over_ranks(time.perf_counter() - t0, dev)
        fl = FLOP_PER_SAMPLE[args.student]
        line["distill"] = dict(metric="distill_samples_per_sec", value=float(nd) * Kd * world / dsec, unit="samples/s", steps=Kd,
                               ms_per_step=1e3 * dsec / Kd, workload="config4 shard: %d envs/GPU, student %s, KL(s||t), TF-Adam, %s"
                               % (nd, args.student, ("gradient all-reduce fused into the student kernel (NVLink peer memory)" if tr.fused_allreduce
                                                     else "NCCL all-reduce of flat grad") if world > 1 else "single GPU"),
                               e2e=dict(value=float(nd) * Kd * world / de2e, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=8,
                                        api="DaggerTrainer.step + wait_loss (rb_dagger_step, rb_dagger_wait_loss: {loss, iteration} posted into mapped host memory)"),
                               gpu_launches_per_step=(2 if tr.student_mode == MODE_TC and (world == 1 or tr.fused_allreduce) else 8),
                               cuda_graph=bool(tr.use_graph), student_mode=("tc" if tr.student_mode == MODE_TC else "fp32"),
                               last_loss=float(tr.last_loss()),
                               roofline=dict(bound="tensor", achieved=fl * nd / (ksec / 20) / 1e12, peak=pk["bf16_burst"], unit="TFLOP/s",
                                             frac=fl * nd / (ksec / 20) / 1e12 / pk["bf16_burst"],
                                             traffic=(STUDENT_TRAFFIC_NCU if (tr.student_mode == MODE_TC and nd == 32768 and args.student == "mlp") else None),
                                             peak_source=pk["src"],
                                             kernel=("k_student_tc (cooperative: fold + tiles + grid reduce + un-fold)" if tr.student_mode == MODE_TC
                                                     else "k_student(loss_grad) + k_reduce_partials"), kernel_ms=1e3 * ksec / 20,
                                             note="tile GEMMs run bf16x3 (3 MMAs per product): tensor-pipe work is 3x the algorithmic FLOP"))
        if tr.student_mode == MODE_TC:                                  # phase stamps (CTA 0, globaltimer) of the last cooperative launch on rank 0
            import ctypes
            tb = (ctypes.c_ulonglong * 48)()
            L.rb_debug_student_timers(tb)
            us = lambda i, j: round((tb[j] - tb[i]) / 1e3, 2)
            ph = dict(fold_image=us(0, 1), sync=us(1, 2), image_load=us(2, 3), tiles=us(3, 4), dump=us(4, 5), sync2=us(5, 6), reduce=us(6, 7), sync3=us(7, 8),
                      unfold=us(8, 32), sync_adam_teardown=us(9, 11), total=us(0, 11))
            if world > 1:
                ph.update(exch_sync=us(32, 33), exch_push=us(33, 34), exch_recv_sum_adam=us(34, 9))
            line["distill"]["student_kernel_phases_us"] = ph
        if world > 1:                                                   # every rank must hold bit-identical student parameters
            chk = torch.stack([tr.student.params.double().sum(), tr.student.params.double().abs().sum()]).to(dev)
            allc = [torch.empty_like(chk) for _ in range(world)]
            torch.distributed.all_gather(allc, chk)
            line["distill"]["ranks_bit_identical"] = bool(all(torch.equal(a, allc[0]) for a in allc))
        line["distill"]["exchange"] = ("none (single rank)" if world == 1 else
                                       "one-shot all-reduce over NVLink peer memory fused into k_student_tc" if tr.fused_allreduce else "NCCL all-reduce")
        tr.close()
        if world > 1:
            # same loop with the gradient exchange done by NCCL (all-reduce kernel + separate Adam launch): the baseline of the fused kernel
            trn = DaggerTrainer(num_envs=nd, seed=0, device=local, student_kind=kind, mode=mode, env_offset=rank * nd, fused_allreduce=False)
            trn.sync_params()
            for _ in range(W):
                trn.step()
            nsec, _ = timed(trn.step, Kd)
            line["distill"]["nccl_baseline"] = dict(value=float(nd) * Kd * world / nsec, unit="samples/s", ms_per_step=1e3 * nsec / Kd)
            trn.close()
        if world == 1:
            # the whole config-4 batch (262 144 envs) on ONE GPU: fixed phases of the cooperative kernel amortise over 14 tiles per SM
            nl = 8 * DISTILL_ENVS_PER_GPU
            trl = DaggerTrainer(num_envs=nl, seed=0, device=local, student_kind=kind, mode=mode)
            for _ in range(W):
                trl.step()
            lsec, _ = timed(trl.step, 30)
            lgl = lambda: trl.student.loss_grad(trl.x, trl.t_pd, s_out=trl.s_pd)
            lgl()
            lksec, _ = timed(lgl, 10)
            line["distill"]["full_batch_one_gpu"] = dict(envs=nl, value=float(nl) * 30 / lsec, unit="samples/s", ms_per_step=1e3 * lsec / 30,
                                                         kernel_ms=1e3 * lksec / 10, tensor_tflops=fl * nl / (lksec / 10) / 1e12,
                                                         frac_of_bf16_peak=fl * nl / (lksec / 10) / 1e12 / pk["bf16_burst"])
            trl.close()
        if world == 1:
            # ---- LSTM student (the reference's headline experiment): one optimiser step on a batch of 10-step windows ---------------
            from reacherdistilation_b200.student_nn import StudentLSTM
            Bw = 2048
            lnet = StudentLSTM(seed=1, device=local, max_batch=Bw)
            lob, lpp = torch.randn((10, Bw, 11), device=dev), torch.randn((10, Bw, 4), device=dev) * 0.3
            ltp = torch.cat([torch.randn((10, Bw, 2), device=dev) * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device=dev)], -1)
            def lstep():
                lnet.step(lob, lpp, ltp, None, keep_prob=0.5, seed=0)      # one CUDA-graph launch (~250 kernels)
            for _ in range(3):
                lstep()
            lsec2, _ = timed(lstep, 10)
            lflop = 6.0 * (243 * 800 + 31400 + 128) * 10 * Bw                       # 3 x 2 x MAC per window row, T = 10
            line["lstm"] = dict(metric="lstm_window_rows_per_sec", value=10.0 * Bw * 10 / lsec2, unit="sample-steps/s", windows=Bw, steps_unrolled=10,
                                ms_per_step=1e3 * lsec2 / 10, tensor_tflops=lflop / (lsec2 / 10) / 1e12, params=int(lnet.P),
                                note="forward + KL + BPTT + Adam of the LSTM(200) student with per-step heads: 34 launches (two persistent tcgen05 recurrence kernels, batched k_gemm_bf16x3 over the un-shared heads) captured once in a CUDA graph (rb_lstm_step, device-side step clock)"
                                     "un-shared heads, element-wise kernels) captured once in a CUDA graph (rb_lstm_step, device-side step clock)")
            del lnet
        # ---- step API: HBM-bound single-step kernel at 4M envs ---------------------------------------------------
        ns = STEP_API_ENVS
        env2 = VecReacher(num_envs=ns, seed=0, device=local, env_offset=0)
        env2.reset()
        act = (torch.rand((ns, 2), device=dev) * 2 - 1)
        sfn = lambda: env2.step(act)
        for _ in range(W):
            sfn()
        ssec, _ = timed(sfn, 50)
        line["step_api"] = dict(metric="reacher_env_steps_per_sec", value=float(ns) * 50 * world / ssec, unit="env-steps/s", envs_per_gpu=ns,
                                roofline=dict(bound="hbm", achieved=ALG_BYTES_STEP * ns / (ssec / 50) / 1e9, peak=pk["hbm"], unit="GB/s",
                                              frac=ALG_BYTES_STEP * ns / (ssec / 50) / 1e9 / pk["hbm"], traffic=(STEP_TRAFFIC_NCU if ns == (1 << 22) else None),
                                              peak_source=pk["src"], kernel="k_step",
                                              note="working set %.0f MB > L2; ncu: 122.9 B of DRAM traffic per env-step (algorithmic 113), issue slots 82 %% "
                                                   "active: the RK4 physics (~750 warp-instructions per step) bounds this kernel before HBM does" % (ns * 97 / 1e6)))
        env2.close()
        if rank == 0 and world == 1:
            cb, _, _ = cpu_reference_leg()
            line["cpu_baseline"] = cb
            line["distill"]["cpu_baseline"] = cpu_distill_leg(args.student)
            line["config1"] = config1_leg()
    if sampler:
        sampler.stop()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
