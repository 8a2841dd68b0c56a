"""Small run of every hot-path kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reacherdistilation_b200 import MODE_TC, MODE_FP32, STUDENT_MLP, STUDENT_POLICY64
from reacherdistilation_b200.env import VecReacher
from reacherdistilation_b200.mlp_train import DaggerTrainer
from reacherdistilation_b200.teacher import init_policy_params
from reacherdistilation_b200.dataset import Dataset
p = torch.from_numpy(init_policy_params(seed=0)).cuda()
env = VecReacher(num_envs=300, seed=0); env.reset()
env.rollout_policy(p, 3, mode=MODE_TC); env.rollout_policy(p, 3, mode=MODE_FP32); env.rollout_random(3)
env.step(torch.zeros((300, 2), device="cuda")); env.close()
for kind in (STUDENT_MLP, STUDENT_POLICY64):
    tr = DaggerTrainer(num_envs=300, seed=1, student_kind=kind, mode=MODE_TC)
    tr.step(); tr.step(); tr.student.forward(tr.x); torch.cuda.synchronize(); tr.close()
ds = Dataset(num_envs=7, generations=2)
for k in range(50): ds.write(torch.randn((7, 11), device="cuda"), torch.randn(7, device="cuda"), torch.randn((7, 4), device="cuda"), None, "t")
ds.flush(); ds.training_batch(5, 10); ds.test_batch(torch.randn((7, 11), device="cuda")); torch.cuda.synchronize(); ds.close()
print("sanitize run done")
