"""rb_gemm_bf16x3 timed with CUDA events (warm L2) on the LSTM's shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200._lib import check, lib, ptr, stream_ptr

def run(M, N, K, a_mn, b_mn, ws_floats, tag):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda"); B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    C = torch.empty((M, N), device="cuda"); ws = torch.empty(max(ws_floats, 1), device="cuda")
    f = lambda: check(lib().rb_gemm_bf16x3(ptr(A), A.shape[1], a_mn, ptr(B), B.shape[1], b_mn, ptr(C), N, M, N, K, None, 0, 0, None, 0,
                                           ptr(ws) if ws_floats else None, ws_floats, stream_ptr()))
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print("%-34s M=%5d N=%4d K=%5d  %7.1f us  %6.1f TFLOP/s" % (tag, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)

for M in (128, 2048, 8192):
    run(M, 800, 243, 0, 1, 0, "fwd recurrence")
    run(M, 243, 800, 0, 0, 0, "bptt dgrad, no split")
    run(M, 243, 800, 0, 0, 32 * M * 243, "bptt dgrad, split-K")
run(243, 800, 20480, 1, 1, 32 * 243 * 800, "wgrad W_l, split-K")
run(2048, 64, 200, 0, 1, 0, "head layer 1")
run(16384, 4096, 4096, 0, 1, 0, "large square-ish")
