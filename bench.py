#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode tc|fp32]
  N > 1: launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Headline workload = BASELINE.json config 3 per GPU: 65 536 lock-step Reacher-v2 envs, fused teacher MLP (2x64 tanh) in the
loop, device-resident rollout buffer; one bench "step" = 20 launches of the 50-step chunk (one episode of every env each) =
65 536 000 env-steps per GPU.  Weak scaling: every rank owns 65 536 envs (global env ids rank*65536 ..), no data-path collective.
`value`   env-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     same metric through the host-buffer C-ABI call (rb_env_rollout_policy_host_ex): H2D of the teacher parameters and D2H of
          the step's result -- the reward of EVERY env-step [T,N] f32 + the per-env done mask [N] u64 (bit t = episode ended at step t) --
          inside the timed region; the rollout buffer (obs, pdflat) is written and stays on the device.  `e2e.done_as_u8`: the same with
          done as [T,N] bytes (round-1 format).  `e2e_episode_returns`: a SECOND figure with the result reduced on the device to per-episode
          returns + done mask.  `e2e_full_buffer`: the same call bringing the whole buffer to the host (PCIe-bound).
`distill` BASELINE.json config 4 shard (32 768 envs per GPU): DAgger iterations = env step + teacher label + student
          forward/backward + KL + [NCCL all-reduce of the flat gradient] + Adam; samples/s == env-steps/s of that loop.
`step_api` the gym-style single-step kernel (HBM-bound) at 4 194 304 envs.
`config1`  BASELINE.json config 1 (one env, 1000-step teacher rollout + one student epoch over those samples): host-thread restatement vs the gym surface at batch 1 vs one fused launch; student epoch in batches of 200.
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
CHUNKS_PER_STEP = 20          # one bench step = 20 launches of the 50-step chunk (1000 steps of every env): `--steps 20` times >= 100 ms
DISTILL_ENVS_PER_GPU = 32768
STEP_API_ENVS = 1 << 22
ALG_BYTES_ROLLOUT = 65.0      # fused rollout writes one buffer row per env-step: ob 44 + pdflat 16 + rew 4 + done 1 B (17 scalars)
MUFU_PER_ENV_STEP = 161.0       # 128 tanh x 1.25 (ex2 each, one rcp per four) + 1 sqrt; the physics no longer uses a reciprocal (csrc/physics.cuh)
XU_LANES_PER_CLK_PER_SM = 16.0  # B200 MUFU rate
ALG_BYTES_STEP = 113.0        # single-step API, I/O 57 B + state round trip 56 B
# ncu counters of the dominant kernels come from the committed captures under profiles/ (never typed in here): see ncu_counters()
NCU_FILES = dict(rollout="profiles/r02_ncu_counts_k_rollout_policy_tc.csv", step="profiles/r02_ncu_counts_k_step.csv",
                 student="profiles/r02_ncu_counts_k_student_tc.csv")
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
FLOP_PER_ENV_STEP = 450.0 + 9856.0          # physics + teacher MLP
FLOP_PER_SAMPLE = {"mlp": 144.4e3, "policy64": 30.3e3}


def ncu_counters(which):
    """Per-launch counters of a kernel from the committed ncu capture (`--csv` long format: one row per launch and metric; averaged over the
    captured launches).  Returns {} when the file is missing: the bench then prints null for those fields instead of inventing them."""
    import csv
    path = os.path.join(ROOT, NCU_FILES[which])
    if not os.path.exists(path):
        return {}
    acc = {}
    with open(path) as fh:
        rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
    if not rows or "Metric Name" not in rows[0]:
        return {}
    iname, ival, iunit = rows[0].index("Metric Name"), rows[0].index("Metric Value"), rows[0].index("Metric Unit")
    for r in rows[1:]:
        try:
            v = float(r[ival].replace(",", ""))
        except ValueError:
            continue
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6}.get(r[iunit], 1.0)       # bytes / ns
        acc.setdefault(r[iname], []).append(v)
    out = {k: sum(v) / len(v) for k, v in acc.items()}
    out["file"] = NCU_FILES[which]
    return out


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


def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not decide the CPU baseline's width)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def teacher_params():
    from reacherdistilation_b200.teacher import init_policy_params      # numpy only: maps no CUDA library
    return init_policy_params(seed=0)


def cpu_rollout(n, chunks, nthreads, p, first_chunk_reward_envs=0):
    """`chunks` teacher-in-the-loop 50-step chunks of n envs (global env ids 0..n-1, seed 0, from reset) on the float64 C restatement.
    Returns (seconds, mean reward of the first chunk over the first `first_chunk_reward_envs` envs)."""
    from oracle import reacher_c as RC
    env = RC.ReacherOracleC(n, seed=0, nthreads=nthreads); env.reset()
    t0 = time.perf_counter()
    first = None
    for c in range(chunks):
        _ob, _pd, rw, _dn, _ = env.rollout_policy(CHUNK_T, p, record=True)
        if c == 0 and first_chunk_reward_envs:
            first = float(rw[:, :first_chunk_reward_envs].mean())
    return time.perf_counter() - t0, first


def cpu_reference_leg(seconds=12.0):
    """Time the C restatement of the reference's CPU rollout (teacher in the loop) on ALL host cores, bounded sample."""
    from oracle import reacher_c as RC
    p, cores = teacher_params(), host_cores()
    cpu_rollout(2048, 1, cores, p)                                       # warm-up (library load, thread pool)
    dt, _ = cpu_rollout(8192, 2, cores, p)
    rate = 8192 * 2 * CHUNK_T / dt
    n = int(min(ENVS_PER_GPU, max(2048, rate * seconds / CHUNK_T)))
    reps = max(1, int(round(seconds * rate / (n * CHUNK_T))))
    dt, _ = cpu_rollout(n, reps, cores, p)
    return dict(value=n * CHUNK_T * reps / dt, unit="env-steps/s", cores=cores, kind="port",
                sample="%d x (%d envs x %d steps) teacher-in-the-loop rollout chunks, float64 C restatement (oracle/reacher_oracle.c, OpenMP %d threads), %.1f s"
                       % (reps, n, CHUNK_T, cores, dt))


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
    return dict(value=B * n_it / dt, unit="samples/s", cores=host_cores(), kind="port",
                sample="%d optimiser steps of batch %d, numpy float64 restatement (oracle/nn_np.py), %.1f s" % (n_it, B, dt))


def config1_leg(steps=1000):
    """BASELINE config 1, the reference's own CPU-runnable case (SURVEY 8(d).1 / CPU plan (i)): ONE env, a 1000-step teacher rollout in the
    per-step loop shape of mlp_train.py:120-139 -- (cpu) the restatement on one host thread, (gpu_gym_loop) the drop-in gym surface at batch 1
    (make_mujoco_env(...).step() + TeacherAgent.mean_and_flat(): one round trip per step to the resident env server, csrc/serve.cu), (gpu_fused) the same 1000 steps as ONE
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
        for _ in range(steps):                              # `ac, _ = pi.act(False, ob); ob, reward, new, _ = env.step(ac)` (mlp_train.py:123-135)
            mean, _flat = teacher.mean_and_flat(ob)
            ob, r, new, _ = env.step(mean)
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
    # ---- the second half of config 1: one student epoch over those `steps` samples as flat batches of 200 (dataset.py:186-194's
    # ratio), student_mlp_graph + KL + Adam per batch: numpy float64 restatement on one host thread vs rb_student_step (one launch per batch)
    from reacherdistilation_b200 import MODE_TC, STUDENT_MLP
    from reacherdistilation_b200.student_nn import StudentNet
    rng = np.random.default_rng(0)
    Bs, nb = 200, steps // 200
    xs, ts = rng.standard_normal((nb, Bs, 16)).astype(np.float32), (rng.standard_normal((nb, Bs, 4)) * 0.3).astype(np.float32)
    for rep in range(2):                                   # first pass = warm-up (BLAS start-up)
        theta, opt = (np.random.default_rng(1).standard_normal(NN.mlp_param_count()) * 0.1), NN.AdamTF(NN.mlp_param_count())
        t0 = time.perf_counter()
        for b in range(nb):
            s_, hs = NN.mlp_fwd(xs[b].astype(np.float64), theta.astype(np.float32)); _l, ds = NN.kl_loss(s_, ts[b].astype(np.float64))
            theta = opt.update(theta, NN.mlp_bwd(hs, theta.astype(np.float32), ds))
        cpu_epoch = nb * Bs / (time.perf_counter() - t0)
    gpu_epoch = {}
    for name, mode in (("tc", MODE_TC), ("fp32", MODE_FP32)):
        net = StudentNet(kind=STUDENT_MLP, seed=1, mode=mode)
        dx, dt_ = torch.from_numpy(xs).cuda(), torch.from_numpy(ts).cuda()
        for rep in range(2):                               # first pass = warm-up
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for b in range(nb):
                net.step(dx[b], dt_[b])
            loss = float(net.gradloss[net.P])              # device -> host read of the epoch's last loss
            gpu_epoch[name] = nb * Bs / (time.perf_counter() - t0)
    return dict(workload="config1: 1 env, %d-step teacher rollout (20 episodes) + one student epoch over those samples (batches of 200)" % steps,
                unit="env-steps/s", cpu_restatement_1_thread=cpu, gpu_gym_loop_batch1=gym_loop, gpu_fused_one_launch=fused,
                teacher_return_per_episode=ret / (steps / 50.0),
                student_epoch=dict(unit="samples/s", batch=Bs, batches=nb, cpu_restatement_numpy=cpu_epoch, gpu_tc=gpu_epoch["tc"], gpu_fp32=gpu_epoch["fp32"],
                                   note="gpu_tc = rb_student_step RB_MODE_TC (one cooperative tcgen05 launch per batch); gpu_fp32 = the fp32 CUDA-core "
                                        "validation path (128-sample tiles: two CTAs at this batch, not built for it)"))


REWARD_CHECK_ENVS = 1024      # both arms print the mean reward of the first 50-step chunk after reset over envs 0..1023 (cross-check)


def workload_config(n):
    return dict(workload="config3: %d envs/GPU, fused teacher MLP (11-64-64-2 tanh) in the loop, rollout buffer device-resident; one step = %d "
                         "chunks of %d env steps" % (n, CHUNKS_PER_STEP, CHUNK_T), envs_per_gpu=n, chunk_steps=CHUNK_T, chunks_per_step=CHUNKS_PER_STEP)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restated in float64 C, oracle/; the reference itself is Python over
    TF-1.10 / gym / MuJoCo-1.50 and cannot be installed here) on ALL host cores.  Each step = the bench step (CHUNKS_PER_STEP chunks) on a
    bounded sample of the envs, sized so that the whole run takes a couple of minutes.  Maps no library of the repo but oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    p, cores = teacher_params(), host_cores()
    _, rew0 = cpu_rollout(2048, 1, cores, p, REWARD_CHECK_ENVS)          # also warms the library and the thread pool up
    dt, _ = cpu_rollout(8192, 2, cores, p)
    rate = 8192 * 2 * CHUNK_T / dt
    per_step_s = min(6.0, 150.0 / (steps + warm))                        # whole run within a few minutes
    n = int(min(ENVS_PER_GPU, max(1024, rate * per_step_s / (CHUNK_T * CHUNKS_PER_STEP))))
    from oracle import reacher_c as RC
    env = RC.ReacherOracleC(n, seed=0, nthreads=cores); env.reset()
    def step():
        for _ in range(CHUNKS_PER_STEP):
            env.rollout_policy(CHUNK_T, p, record=True)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = n * CHUNK_T * CHUNKS_PER_STEP * steps / dt
    sample = ("each step = %d of the %d envs x %d chunks x %d steps, float64 C restatement of the MuJoCo + TF CPU path (oracle/reacher_oracle.c), "
              "OpenMP %d threads" % (n, ENVS_PER_GPU, CHUNKS_PER_STEP, CHUNK_T, cores))
    cfg = workload_config(ENVS_PER_GPU)
    cfg["reference_sample"] = sample
    cfg["mean_reward_first_chunk_envs_0_%d" % REWARD_CHECK_ENVS] = rew0
    line = dict(impl="reference", metric="reacher_env_steps_per_sec", value=val, unit="env-steps/s", n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=cfg,
                cpu_baseline=dict(value=val, unit="env-steps/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="env-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def distill_parity(DaggerTrainer, kind, mode, nd, rank, world, local, dev, iters=3):
    """Data-parallel parity of the student step at this world size, from a fixed seed (MpiAdam.update = Allreduce + Adam,
    backup/student_rollout.py:658-659,709):  (a) `iters` DAgger iterations with the exchange fused into k_student_tc vs the same with NCCL
    all-reduce + rb_adam_step;  (b) vs ONE GPU running the concatenated batch (world x nd envs, same global env ids => same samples and
    dropout masks; KL is a sum, so the summed shard gradients are the gradient of the concatenated batch up to fp32 summation order)."""
    import torch
    out = dict(iterations=iters, world=world, envs_per_rank=nd)

    def run(**kw):
        tr = DaggerTrainer(num_envs=kw.pop("n", nd), seed=0, device=local, student_kind=kind, mode=mode, lr=1e-3, **kw)
        tr.sync_params()
        losses = []
        for _ in range(iters):
            tr.step()
            losses.append(float(tr.last_loss()))
        p = tr.student.params.double().cpu()
        tr.close()
        return losses, p
    lf, pf = run(env_offset=rank * nd)                                    # fused exchange (default at world > 1 in tc mode)
    ln, pn = run(env_offset=rank * nd, fused_allreduce=False)             # NCCL
    rel = lambda a, b: max(abs(x - y) / max(1.0, abs(y)) for x, y in zip(a, b))
    out["fused_vs_nccl"] = dict(loss_rel=rel(lf, ln), param_abs=float((pf - pn).abs().max()), tol=dict(loss_rel=1e-5, param_abs=1e-5))
    torch.distributed.barrier()
    if rank == 0:                                                         # a single-rank trainer inside the multi-rank job; the others wait below
        tr = DaggerTrainer(num_envs=world * nd, seed=0, device=local, student_kind=kind, mode=mode, lr=1e-3, env_offset=0, solo=True)
        l1 = []
        for _ in range(iters):
            tr.step()
            l1.append(float(tr.last_loss()))
        p1 = tr.student.params.double().cpu()
        tr.close()
        scale = float(max(1.0, p1.abs().max()))
        out["fused_vs_single_gpu_concatenated_batch"] = dict(loss_rel=rel(lf, l1), param_abs=float((pf - p1).abs().max()), param_scale=scale,
                                                              tol=dict(loss_rel=2e-6, param_abs=2e-5 * scale), envs=world * nd)
    torch.distributed.barrier()
    ok = out["fused_vs_nccl"]["loss_rel"] <= 1e-5 and out["fused_vs_nccl"]["param_abs"] <= 1e-5
    if rank == 0:
        c = out["fused_vs_single_gpu_concatenated_batch"]
        ok = ok and c["loss_rel"] <= c["tol"]["loss_rel"] and c["param_abs"] <= c["tol"]["param_abs"]
    flag = torch.tensor([1 if ok else 0], device=dev)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    return out


def run_sweep(args):
    """--sweep: BASELINE.json config 5 -- total envs 2^10 .. 2^22 over the launched GPUs (contiguous global-id shards, weak per-size work split):
    fused teacher rollout and single-step API, CUDA-event time, max over ranks; on rank 0 the host-core restatement at EVERY size (bounded)."""
    import torch
    from reacherdistilation_b200 import MODE_FP32, MODE_TC, _lib
    from reacherdistilation_b200.dist import init_from_env, max_over_ranks, shard_range
    from reacherdistilation_b200.env import VecReacher
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    mode = MODE_TC if L.rb_mode_available(MODE_TC) else MODE_FP32
    p_host = teacher_params()
    teacher = torch.from_numpy(p_host).to(dev)
    cores = host_cores()
    rows = []

    def timed(fn, iters):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    for lg in range(10, 23):
        total = 1 << lg
        lo, hi = shard_range(total, rank, world)
        n = hi - lo
        env = VecReacher(num_envs=max(n, 1), seed=0, device=local, env_offset=lo)
        env.reset()
        buf = dict(obs=torch.empty((CHUNK_T, max(n, 1), 11), device=dev), pdflat=torch.empty((CHUNK_T, max(n, 1), 4), device=dev),
                   rew=torch.empty((CHUNK_T, max(n, 1)), device=dev), done=torch.empty((CHUNK_T, max(n, 1)), dtype=torch.uint8, device=dev))
        roll = lambda: env.rollout_policy(teacher, CHUNK_T, nout=2, mode=mode, out=buf)
        for _ in range(3):
            roll()
        it = 20 if lg <= 18 else 6
        rsec = timed(roll, it)
        act = torch.rand((max(n, 1), 2), device=dev) * 2 - 1
        stp = lambda: env.step(act)
        for _ in range(3):
            stp()
        ssec = timed(stp, 50)
        env.close()
        del buf
        row = dict(total_envs=total, envs_per_gpu=n, rollout_env_steps_per_s=float(total) * CHUNK_T * it / rsec, rollout_ms_per_chunk=1e3 * rsec / it,
                   step_api_env_steps_per_s=float(total) * 50 / ssec)
        if rank == 0:                                                    # host-core baseline at every size: ~1.5 s each
            nc = min(total, 65536)
            dt, _ = cpu_rollout(nc, 1, cores, p_host)
            reps = max(1, int(1.5 / max(dt, 1e-3)))
            dt, _ = cpu_rollout(nc, reps, cores, p_host)
            row["cpu_env_steps_per_s"] = nc * CHUNK_T * reps / dt
            row["cpu_sample"] = "%d x (%d envs x %d steps)" % (reps, nc, CHUNK_T)
        rows.append(row)
    if rank == 0:
        print(json.dumps(dict(metric="reacher_env_steps_per_sec", sweep="config5: total envs 2^10..2^22 over %d GPU(s)" % world, n_gpus=world, unit="env-steps/s",
                              scaling="strong (fixed total per row, sharded by contiguous global env id)", cpu_cores=cores,
                              cpu_kind="port (float64 C restatement, OpenMP)", rows=rows)))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fp32", "tc"])
    ap.add_argument("--student", default="mlp", choices=["mlp", "policy64"])
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements (distill, step API, CPU legs)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE config 5: 1 k .. 4 M envs on the launched GPUs + host-core baseline at every size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.sweep:
        return run_sweep(args)

    import numpy as np
    import torch
    from reacherdistilation_b200 import MODE_FP32, MODE_TC, STUDENT_MLP, STUDENT_POLICY64, _lib
    from reacherdistilation_b200.dist import init_from_env, max_over_ranks
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    from reacherdistilation_b200.teacher import init_policy_params
    from reacherdistilation_b200._lib import check

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
    n, R = ENVS_PER_GPU, CHUNKS_PER_STEP
    teacher = torch.from_numpy(init_policy_params(seed=0)).to(dev)
    env = VecReacher(num_envs=n, seed=0, device=local, env_offset=rank * n)
    env.reset()
    buf = dict(obs=torch.empty((CHUNK_T, n, 11), device=dev), pdflat=torch.empty((CHUNK_T, n, 4), device=dev),
               rew=torch.empty((CHUNK_T, n), device=dev), done=torch.empty((CHUNK_T, n), dtype=torch.uint8, device=dev))
    chunk_fn = lambda: env.rollout_policy(teacher, CHUNK_T, nout=2, mode=mode, out=buf)

    def step_fn():                                                     # one bench step = R chunk launches (1000 env steps of every env)
        for _ in range(R):
            chunk_fn()
    chunk_fn()                                                         # first chunk after reset: the reward both arms print (cross-check)
    rew_check = float(buf["rew"][:, :REWARD_CHECK_ENVS].mean()) if rank == 0 else None
    for _ in range(W):
        step_fn()
    sec, win = timed(step_fn, K)
    env_steps = float(n) * CHUNK_T * R * K * world
    value = env_steps / sec
    kernel_s = sec / (K * R)                                           # one kernel launch per chunk
    clocks = sampler.window(*win) if sampler else None
    mean_rew = float(buf["rew"].mean())

    # ---- e2e: the host-buffer C-ABI call (rb_env_rollout_policy_host), synchronous, copies inside the timed region -------------
    #   e2e             : H2D of the teacher parameters, the rollout (buffer stays device-resident for the distillation loop, north_star (3)),
    #                     D2H of the step's RESULT: reward and done of every env-step (what the reference prints / extract_reward.py consumes)
    #   e2e_full_buffer : same call returning the WHOLE rollout buffer (obs, pdflat, reward, done = 65 B/env-step) to the host -- PCIe-bound
    Ke = max(3, min(K, 20))
    hbuf = dict(obs=torch.empty((CHUNK_T, n, 11)).pin_memory(), pdflat=torch.empty((CHUNK_T, n, 4)).pin_memory(),
                rew=torch.empty((CHUNK_T, n)).pin_memory(), done=torch.empty((CHUNK_T, n), dtype=torch.uint8).pin_memory())
    hmask = torch.empty((n,), dtype=torch.int64).pin_memory()
    hres = dict(obs=None, pdflat=None, rew=hbuf["rew"], done=None, done_mask=hmask)         # the step's RESULT: reward of every env-step + the done mask
    hres_u8 = dict(obs=None, pdflat=None, rew=hbuf["rew"], done=hbuf["done"])               # same with `done` as T bytes per env (copy engine, slab by slab)
    tparams_host = torch.from_numpy(init_policy_params(seed=0)).pin_memory()

    def e2e_time(out, chunks):
        fn = lambda: env.rollout_policy_host(tparams_host, CHUNK_T, nout=2, mode=mode, out=out)
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke * chunks):
            fn()                                                       # synchronous call (copies + sync inside)
        barrier()
        return max_over_ranks(time.perf_counter() - t0, dev)
    def e2e_time_pipelined(chunks):
        # the split-phase form of the same call (rb_env_rollout_policy_host_begin / _wait), two calls in flight on alternating result buffers:
        # every call still copies its parameters in and delivers its result to the host; the host-side work around call i overlaps kernel i + 1
        outs = [hres, dict(rew=torch.empty((CHUNK_T, n)).pin_memory(), done_mask=torch.empty((n,), dtype=torch.int64).pin_memory())]
        for o in outs:
            env.rollout_policy_host_begin(tparams_host, CHUNK_T, nout=2, mode=mode, out=o); env.rollout_policy_host_wait()
        barrier()
        t0 = time.perf_counter()
        total = Ke * chunks
        for i in range(total):
            env.rollout_policy_host_begin(tparams_host, CHUNK_T, nout=2, mode=mode, out=outs[i & 1])
            if i >= 1:
                env.rollout_policy_host_wait()
        env.rollout_policy_host_wait()
        barrier()
        return max_over_ranks(time.perf_counter() - t0, dev)
    hret = torch.empty((n,), dtype=torch.float32).pin_memory()
    hres_ret = dict(obs=None, pdflat=None, rew=None, done=None, done_mask=hmask, return_sum=hret)   # per-episode returns reduced on the device
    e2e_sync_sec, e2e_u8_sec, e2e_ret_sec, e2e_full_sec = e2e_time(hres, R), e2e_time(hres_u8, R), e2e_time(hres_ret, R), e2e_time(hbuf, 1)
    e2e_ks_sec = e2e_time_pipelined(R)                      # reward stored by the kernel into the host buffer (default transport)
    check(L.rb_env_set_host_transport(env._h, 0))           # reward through a per-call device buffer + the copy engine, beside the next call's kernel
    e2e_ce_sec = e2e_time_pipelined(R)
    check(L.rb_env_set_host_transport(env._h, 1))
    # transport rule (measured, profiles/README.md): one GPU per host link -> kernel-posted stores (1.05e10 vs 1.02e10); several ranks sharing the
    # host's PCIe ingest -> copy engine (2 ranks: 1.71e10 vs 1.58e10; the posted 128-byte stores of several GPUs contend on the shared uplink)
    e2e_sec = e2e_ks_sec if world == 1 else e2e_ce_sec
    e2e = dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_sec, unit="env-steps/s", h2d_bytes_per_step=int(tparams_host.numel() * 4) * R,
               d2h_bytes_per_step=int(n * CHUNK_T * 4 + n * 8) * R, steps=Ke,
               api="rb_env_rollout_policy_host_begin / _wait (VecReacher.rollout_policy_host_begin / _wait), %d calls per step, two in flight" % R,
               transport="kernel-posted reward stores (rb_env_set_host_transport 1)" if world == 1 else "copy engine for the reward (rb_env_set_host_transport 0)",
               split_phase_kernel_stored_reward=dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_ks_sec),
               split_phase_copy_engine_reward=dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_ce_sec,
                                                   note="rb_env_set_host_transport 0: the reward of call i goes through a device buffer and the copy engine while kernel i + 1 runs"),
               synchronous_call=dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_sync_sec, api="rb_env_rollout_policy_host_ex: one call at a time, "
                                     "returns when the result is in host memory"),
               result="reward[T,N] f32 of every env-step + done_mask[N] u64 (bit t = the env finished an episode at step t of the chunk) of every chunk, "
                      "both written by the kernel into the caller's page-locked buffers (no copy-engine transfer); obs / pdflat are written to the "
                      "device-resident rollout buffer (rb_env_rollout_buffer) and stay there",
               mean_reward_host=float(hbuf["rew"].mean()), episodes_finished_last_chunk=int(sum(bin(int(v) & ((1 << 64) - 1)).count("1") for v in hmask[:4096].tolist())),
               done_as_u8=dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_u8_sec, d2h_bytes_per_step=int(n * CHUNK_T * (4 + 1)) * R,
                               note="round-1 result format: done[T,N] u8 copied per in-kernel progress slab while the launch runs"))
    e2e_ret = dict(value=float(n) * CHUNK_T * R * Ke * world / e2e_ret_sec, unit="env-steps/s", h2d_bytes_per_step=int(tparams_host.numel() * 4) * R,
                   d2h_bytes_per_step=int(n * (4 + 8)) * R, steps=Ke, mean_episode_return_host=float(hret.mean()),
                   note="a SECOND end-to-end figure, not the headline: the step's result reduced on the device to what the reference logs per episode "
                        "(mlp_train.py:129-139) -- return_sum[N] f32 (each 50-step chunk starts at an episode boundary, so it is the episode return) + "
                        "done_mask[N] u64; 12 bytes per env and chunk instead of 208")
    e2e_full = dict(value=float(n) * CHUNK_T * Ke * world / e2e_full_sec, unit="env-steps/s", h2d_bytes_per_step=int(tparams_host.numel() * 4),
                    d2h_bytes_per_step=int(n * CHUNK_T * (44 + 16 + 4 + 1)), steps=Ke, note="ONE chunk per step, whole rollout buffer to the host: PCIe-bound")
    env.close()
    del buf, hbuf

    cfg = workload_config(n)
    cfg.update(policy_mode=mode_name, seed=0, mean_teacher_reward=mean_rew,
               l2="no flush: each chunk writes a fresh %.0f MB rollout buffer (> 126 MB L2); state 3.1 MB stays in registers" % (n * CHUNK_T * 68 / 1e6))
    cfg["mean_reward_first_chunk_envs_0_%d" % REWARD_CHECK_ENVS] = rew_check
    line = dict(metric="reacher_env_steps_per_sec", value=value, unit="env-steps/s", n_gpus=world, steps=K, warmup=W, ms_per_step=1e3 * sec / K,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=cfg,
                e2e=e2e, e2e_episode_returns=e2e_ret, e2e_full_buffer=e2e_full, gpu_launches=K * R, clocks=clocks)
    nc = ncu_counters("rollout") if (mode == MODE_TC and n == 65536) else {}
    warps = n / 32.0 * CHUNK_T
    sms = L.rb_sm_count(local)
    xu_ceiling = sms * XU_LANES_PER_CLK_PER_SM * 1.965e9 / MUFU_PER_ENV_STEP
    line["roofline"] = dict(bound="hbm", achieved=ALG_BYTES_ROLLOUT * n * CHUNK_T / kernel_s / 1e9, peak=pk["hbm"], unit="GB/s",
                            frac=ALG_BYTES_ROLLOUT * n * CHUNK_T / kernel_s / 1e9 / pk["hbm"],
                            traffic=(nc["dram__bytes_read.sum"] + nc["dram__bytes_write.sum"]) if "dram__bytes_read.sum" in nc else None,
                            peak_source=pk["src"], kernel="k_rollout_policy_%s" % ("tc" if mode == MODE_TC else "fp32"), kernel_ms=1e3 * kernel_s,
                            bound_binding="xu+issue",
                            binding=dict(frac_of_xu_ceiling=value / world / xu_ceiling, xu_ceiling_env_steps_per_s=xu_ceiling,
                                         issue_active_pct=nc.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                         tensor_pipe_active_pct=nc.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                         warp_instructions_per_warp_step=(nc["smsp__inst_executed.sum"] / warps) if "smsp__inst_executed.sum" in nc else None,
                                         xu_warp_instructions_per_warp_step=(nc["sm__inst_executed_pipe_xu.sum"] / warps) if "sm__inst_executed_pipe_xu.sum" in nc else None,
                                         ncu_file=nc.get("file")),
                            note="`frac` is the contract's HBM figure (65 algorithmic bytes per env-step over the measured copy peak); HBM is NOT what binds "
                                 "the fused rollout -- it only writes the buffer row, the state stays in registers (SURVEY 8(d)).  The binding resources are "
                                 "the XU (MUFU) pipe and instruction issue on 13.8 warps per SM: `binding` (ncu counters read from the committed capture)")
    line["pipes"] = dict(xu_mufu_per_env_step=MUFU_PER_ENV_STEP, xu_ceiling_env_steps_per_s=xu_ceiling, frac_of_xu_ceiling=value / world / xu_ceiling,
                         tensor_tflops_bf16x3=3 * 9856.0 * n * CHUNK_T / kernel_s / 1e12, tensor_peak_tflops=pk["bf16_burst"],
                         physics_flop_per_env_step=450.0, teacher_flop_per_env_step=9856.0,
                         note="XU ceiling = SMs x 16 MUFU lanes/clk x 1.965 GHz / 161 MUFU per env-step (128 tanh x 1.25 + 1 sqrt)")

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
        lg = lambda: tr.student.loss_grad(tr.x, tr.t_pd, s_out=tr.s_pd)      # tc mode: weight-image kernel + ONE cooperative launch (tiles, reduce, un-fold)
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
                                             traffic=((lambda c: (c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"]) if "dram__bytes_read.sum" in c else None)(
                                                 ncu_counters("student")) if (tr.student_mode == MODE_TC and nd == 32768 and args.student == "mlp") else None),
                                             peak_source=pk["src"],
                                             kernel=("k_student_image + k_student_tc (cooperative: tiles + grid reduce + un-fold)" if tr.student_mode == MODE_TC
                                                     else "k_student(loss_grad) + k_reduce_partials"), kernel_ms=1e3 * ksec / 20,
                                             note="tile GEMMs run bf16x3 (3 MMAs per product): tensor-pipe work is 3x the algorithmic FLOP"))
        if tr.student_mode == MODE_TC:                                  # phase stamps (CTA 0, globaltimer) of the last cooperative launch on rank 0
            import ctypes
            tb = (ctypes.c_ulonglong * 48)()
            L.rb_debug_student_timers(tb)
            us = lambda i, j: round((tb[j] - tb[i]) / 1e3, 2)
            ph = dict(setup=us(0, 1), image_wait=us(1, 3), tiles=us(3, 4), dump_and_env_step=us(4, 5), sync1=us(5, 6), reduce=us(6, 7), sync2=us(7, 8),
                      teardown=us(9, 11), total=us(0, 11), grid_barriers=2)
            if world > 1:
                ph.update(unfold_push=us(8, 34), recv_sum_adam=us(34, 9))
            else:
                ph.update(unfold_adam=us(8, 9))
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
            line["distill"]["parity"] = distill_parity(DaggerTrainer, kind, mode, nd, rank, world, local, dev)
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
                                note="forward + KL + BPTT + Adam of the LSTM(200) student with per-step heads: 33 launches (two persistent tcgen05 recurrence kernels, "
                                     "batched k_gemm_bf16x3 over the un-shared heads at two / three CTAs per SM, element-wise kernels) captured once in a CUDA graph "
                                     "(rb_lstm_step, device-side step clock)")
            del lnet
            # ---- two-headed LSTM student of the backup experiment (backup/student_rollout.py:130-200,328) at the sizes the source comments
            # beside its debug values (NUM_UNITS 100, STEPS_UNROLLED 20, LSTM_BATCH_SIZE 100) and at 2048 windows: loss_grad + Adam ------------
            from reacherdistilation_b200.student_nn import StudentLSTM2, lstm2_spec
            l2 = {}
            for Bw2 in (100, 2048):
                n2 = StudentLSTM2(spec=lstm2_spec(units=100, steps=20), seed=1, device=local)
                o2, a2 = torch.randn((20, Bw2, 11), device=dev), torch.randn((20, Bw2, 2), device=dev) * 0.3
                t2 = torch.cat([torch.randn((20, Bw2, 2), device=dev) * 0.3, -1 + 0.2 * torch.randn((20, Bw2, 2), device=dev)], -1)
                r2 = torch.randn((20, Bw2), device=dev) * 0.2
                def l2step():
                    n2.step(o2, a2, t2, r2, None, keep_prob=0.5, seed=0)          # one CUDA-graph launch (rb_lstm2_step)
                for _ in range(3):
                    l2step()
                s2, _ = timed(l2step, 10)
                l2["windows_%d" % Bw2] = dict(ms_per_step=1e3 * s2 / 10, value=20.0 * Bw2 * 10 / s2, unit="sample-steps/s")
                l2["params"] = int(n2.P)
                del n2
            line["lstm2"] = dict(metric="lstm2_window_rows_per_sec", graph="source variant: units 100, 20 unrolled steps, state not carried, trunk 128 -> "
                                 "reward 64-1 || action 64-4, loss = KL + squared reward error", **l2)
        # ---- step API: HBM-bound single-step kernel at 4M envs ---------------------------------------------------
        ns = STEP_API_ENVS
        env2 = VecReacher(num_envs=ns, seed=0, device=local, env_offset=0)
        env2.reset()
        # action model of BASELINE config 2: a fresh U(-1, 1) action for every env at every step (8 pre-generated tensors, cycled).  A CONSTANT action
        # per env -- what this leg used before -- drives every arm into its joint limit within 0.3 s and keeps it there; that case is still
        # reported as `pinned_at_joint_limit` (measured: within 1-2 % of the fresh-action figure).
        acts = [(torch.rand((ns, 2), device=dev) * 2 - 1) for _ in range(8)]
        cnt = [0]
        def sfn():
            env2.step(acts[cnt[0] & 7]); cnt[0] += 1
        for _ in range(W + 50):                              # past the first episode boundary: the envs are spread over their state space
            sfn()
        ssec, _ = timed(sfn, 50)
        pin = lambda: env2.step(acts[0])
        for _ in range(60):
            pin()
        psec, _ = timed(pin, 50)
        line["step_api"] = dict(metric="reacher_env_steps_per_sec", value=float(ns) * 50 * world / ssec, unit="env-steps/s", envs_per_gpu=ns,
                                roofline=dict(bound="hbm", achieved=ALG_BYTES_STEP * ns / (ssec / 50) / 1e9, peak=pk["hbm"], unit="GB/s",
                                              frac=ALG_BYTES_STEP * ns / (ssec / 50) / 1e9 / pk["hbm"],
                                              traffic=(lambda c: (c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"]) if ("dram__bytes_read.sum" in c and ns == (1 << 22)) else None)(ncu_counters("step")),
                                              peak_source=pk["src"], kernel="k_step", kernel_ms=1e3 * ssec / 50,
                                              issue_active_pct=ncu_counters("step").get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                              note="working set %.0f MB > L2; a fresh uniform action per env and step (BASELINE config 2's action model); instruction "
                                                   "issue bounds this kernel before HBM does (real DRAM traffic: `traffic`, 96 B of it the two-float state round trip)" % (ns * 153 / 1e6)),
                                pinned_at_joint_limit=dict(value=float(ns) * 50 * world / psec, kernel_ms=1e3 * psec / 50,
                                                           frac=ALG_BYTES_STEP * ns / (psec / 50) / 1e9 / pk["hbm"],
                                                           note="constant action per env: every arm sits on its joint limit, the contact branch runs in all 8 RK4 "
                                                                "stages (the workload the round-1 line and the committed ncu capture used)"))
        env2.close()
        if rank == 0 and world == 1:
            line["cpu_baseline"] = cpu_reference_leg()
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
