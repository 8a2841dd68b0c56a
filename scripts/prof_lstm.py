"""A few LSTM student optimiser steps (2048 windows) without the graph -- the command profiled with ncu (launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200.student_nn import StudentLSTM
Bw = 2048
net = StudentLSTM(seed=1, max_batch=Bw)
ob, pp = torch.randn((10, Bw, 11), device="cuda"), torch.randn((10, Bw, 4), device="cuda") * 0.3
tp = torch.cat([torch.randn((10, Bw, 2), device="cuda") * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device="cuda")], -1)
for _ in range(3):
    net.loss_grad(ob, pp, tp, None, keep_prob=0.5, seed=0, iteration=net.t); net.adam_step()
torch.cuda.synchronize()
print("done")
