"""LSTM student optimiser step (CUDA graph) timed with events at a few window-batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200.student_nn import StudentLSTM
for Bw in (20, 256, 2048, 8192):
    net = StudentLSTM(seed=1, max_batch=Bw)
    ob, pp = torch.randn((10, Bw, 11), device="cuda"), torch.randn((10, Bw, 4), device="cuda") * 0.3
    tp = torch.cat([torch.randn((10, Bw, 2), device="cuda") * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device="cuda")], -1)
    for _ in range(3):
        net.step(ob, pp, tp, None, keep_prob=0.5, seed=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net.step(ob, pp, tp, None, keep_prob=0.5, seed=0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flop = 6.0 * (243 * 800 + 31400 + 128) * 10 * Bw
    print("LSTM step, %5d windows x 10: %.3f ms  %.3e sample-steps/s  %.1f TFLOP/s" % (Bw, ms, 10 * Bw / ms * 1e3, flop / ms / 1e9), flush=True)
    del net
