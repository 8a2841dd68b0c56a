import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from reacherdistilation_b200 import MODE_TC
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.teacher import init_policy_params
n, T = 65536, 50
env = VecReacher(num_envs=n, seed=0); env.reset()
p = torch.from_numpy(init_policy_params(seed=0)).pin_memory()
out = dict(obs=None, pdflat=None, rew=torch.empty((T, n)).pin_memory(), done=torch.empty((T, n), dtype=torch.uint8).pin_memory())
for _ in range(3): env.rollout_policy_host(p, T, nout=2, mode=MODE_TC, out=out)
t0 = time.perf_counter()
for _ in range(48): env.rollout_policy_host(p, T, nout=2, mode=MODE_TC, out=out)
print("ms per call %.4f" % ((time.perf_counter() - t0) / 48 * 1e3))
