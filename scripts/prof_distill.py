"""A few DAgger iterations on the config-4 shard (32 768 envs, MLP student, tcgen05) -- the command profiled with ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200 import MODE_TC, STUDENT_MLP, STUDENT_POLICY64
from reacherdistilation_b200.mlp_train import DaggerTrainer
kind = STUDENT_POLICY64 if "policy64" in sys.argv else STUDENT_MLP
tr = DaggerTrainer(num_envs=32768, seed=0, student_kind=kind, mode=MODE_TC)
for _ in range(8):
    tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tr.step()
e1.record(); torch.cuda.synchronize()
print("dagger step %.1f us, loss %.1f" % (e0.elapsed_time(e1) * 1e3 / 20, float(tr.last_loss())))
