"""Fused teacher rollout: time per 50-step chunk vs number of envs (tiles per SM) -- how well tiles overlap on one SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200 import MODE_TC, MODE_FP32
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.teacher import init_policy_params
p = torch.from_numpy(init_policy_params(seed=0)).cuda()
T = 50
for n in [148 * 32, 148 * 64, 148 * 128, 148 * 256, 148 * 384, 65536, 148 * 512, 148 * 1024, 148 * 2048, 1 << 20]:
    env = VecReacher(num_envs=n, seed=0); env.reset()
    buf = dict(obs=torch.empty((T, n, 11), device="cuda"), pdflat=torch.empty((T, n, 4), device="cuda"), rew=torch.empty((T, n), device="cuda"),
               done=torch.empty((T, n), dtype=torch.uint8, device="cuda"))
    for _ in range(3): env.rollout_policy(p, T, mode=MODE_TC, out=buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): env.rollout_policy(p, T, mode=MODE_TC, out=buf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("envs %8d  warps/SM %6.1f  chunk %.3f ms  step %.2f us  %.3e env-steps/s" % (n, n / 32 / 148, ms, ms * 1e3 / T, n * T / ms * 1e3), flush=True)
    env.close(); del buf
