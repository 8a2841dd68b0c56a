"""BASELINE.json config 5: throughput sweep 1k - 4M envs (TOTAL over the ranks) of the fused teacher rollout and of the single-step API.

  python scripts/sweep_scaling.py                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sweep_scaling.py

Env shards are contiguous global-id ranges (no data-path collective); time = CUDA events, max over ranks; rank 0 prints one JSON line per size
and writes gpurun_out/sweep_<N>gpu.json.  The host-core baseline (float64 C restatement of the MuJoCo path, OpenMP) is timed at every size on
rank 0 of the 1-GPU run on a bounded sample.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200 import MODE_TC
from reacherdistilation_b200.dist import init_from_env, max_over_ranks, shard_range
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.teacher import init_policy_params

rank, world, local = init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
p = torch.from_numpy(init_policy_params(seed=0)).to(dev)
T = 50


def timed(fn, iters):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)


rows = []
for lg in range(10, 23):
    total = 1 << lg
    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    env = VecReacher(num_envs=n, seed=0, device=local, env_offset=lo)
    env.reset()
    buf = dict(obs=torch.empty((T, n, 11), device=dev), pdflat=torch.empty((T, n, 4), device=dev), rew=torch.empty((T, n), device=dev),
               done=torch.empty((T, n), dtype=torch.uint8, device=dev))
    roll = lambda: env.rollout_policy(p, T, nout=2, mode=MODE_TC, out=buf)
    for _ in range(3):
        roll()
    it = 20 if lg <= 18 else 5
    sec = timed(roll, it)
    act = torch.rand((n, 2), device=dev) * 2 - 1
    step = lambda: env.step(act)
    for _ in range(3):
        step()
    ssec = timed(step, 50)
    row = dict(envs_total=total, n_gpus=world, envs_per_gpu=n, rollout_env_steps_per_s=total * T * it / sec, rollout_ms_per_chunk=1e3 * sec / it,
               step_api_env_steps_per_s=total * 50 / ssec, step_api_us=1e6 * ssec / 50)
    if rank == 0 and world == 1:
        from oracle import reacher_c as RC                      # checker timed as the reported CPU baseline (bench leg)
        import numpy as np
        nc = min(total, 1 << 16)
        c = RC.ReacherOracleC(nc, seed=0); c.reset()
        pp = init_policy_params(seed=0)
        t0 = time.perf_counter(); reps = 0
        while reps == 0 or time.perf_counter() - t0 < 1.0:
            c.rollout_policy(T, pp, record=True); reps += 1
        row["cpu_env_steps_per_s"] = nc * T * reps / (time.perf_counter() - t0)
        row["cpu_cores"] = RC.max_threads()
    env.close(); del buf
    if rank == 0:
        print(json.dumps(row), flush=True)
        rows.append(row)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/sweep_%dgpu.json" % world, "w"), indent=1)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
