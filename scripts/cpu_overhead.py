"""Host-side cost of issuing the hot-path calls (ctypes + Python) vs their device time: are the timed loops launch-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200 import MODE_TC, STUDENT_MLP
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.mlp_train import DaggerTrainer

def measure(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    t_issue = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return t_issue / n * 1e6, e0.elapsed_time(e1) * 1e3 / n

tr = DaggerTrainer(num_envs=32768, seed=0, student_kind=STUDENT_MLP, mode=MODE_TC)
print("dagger step      : host issue %.1f us, device %.1f us" % measure(tr.step))
tiny = DaggerTrainer(num_envs=128, seed=0, student_kind=STUDENT_MLP, mode=MODE_TC)
print("dagger step (128): host issue %.1f us, device %.1f us" % measure(tiny.step))
n = 1 << 22
env = VecReacher(num_envs=n, seed=0); env.reset()
act = torch.rand((n, 2), device="cuda") * 2 - 1
print("env.step 4M      : host issue %.1f us, device %.1f us" % measure(lambda: env.step(act), 50))
small = VecReacher(num_envs=1024, seed=0); small.reset()
a2 = torch.zeros((1024, 2), device="cuda")
print("env.step 1k      : host issue %.1f us, device %.1f us" % measure(lambda: small.step(a2)))
