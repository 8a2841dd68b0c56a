"""Single-step API kernel (k_step) at 4 M envs -- the command profiled with ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200.env import VecReacher
n = 1 << 22
env = VecReacher(num_envs=n, seed=0); env.reset()
act = torch.rand((n, 2), device="cuda") * 2 - 1
for _ in range(5): env.step(act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): env.step(act)
e1.record(); torch.cuda.synchronize()
print("k_step %.1f us per step, %.3e env-steps/s" % (e0.elapsed_time(e1) * 1e3 / 20, n * 20 / e0.elapsed_time(e1) * 1e3))
