import torch, sys
sys.path.insert(0, '.')
from reacherdistilation_b200.env import VecReacher
n = 1 << 22
env = VecReacher(num_envs=n, seed=0); env.reset()
act = torch.rand((n, 2), device="cuda") * 2 - 1
def t(k):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(k): env.step(act)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) * 1e3 / k
print("steps 0-4 after reset: %.1f us" % t(5)); print("steps 5-14: %.1f us" % t(10)); print("steps 15-24: %.1f us" % t(10)); print("steps 25-49: %.1f us" % t(25)); print("steps 50-99: %.1f us" % t(50)); print("steps 100-149: %.1f us" % t(50))
