"""A/B of rb_gemm_set_cta_packing (0: deep pipeline, 1: packed CTAs, -1: by grid size) on the LSTM optimiser step (CUDA graph, 2048 windows),
the two-headed LSTM step and a few stand-alone GEMM shapes.  CUDA events on the launching stream, medians of repeated timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from reacherdistilation_b200._lib import check, lib, ptr, stream_ptr
from reacherdistilation_b200.student_nn import StudentLSTM, StudentLSTM2, lstm2_spec

dev = "cuda"


def timed(fn, n, reps=5):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / n)
    return sorted(out)[len(out) // 2]


def gemm(M, N, K, a_mn, b_mn, ws_floats=0):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.empty((M, N), device=dev)
    ws = torch.empty(ws_floats, device=dev) if ws_floats else None
    def run():
        check(lib().rb_gemm_bf16x3(ptr(A), A.shape[1], a_mn, ptr(B), B.shape[1], b_mn, ptr(Cm), N, M, N, K, None, 0, 0, None, 0, ptr(ws), ws_floats, stream_ptr()))
    return run


for mode in (0, 1, -1):
    check(lib().rb_gemm_set_cta_packing(mode))
    Bw = 2048
    net = StudentLSTM(seed=1, max_batch=Bw)
    ob, pp = torch.randn((10, Bw, 11), device=dev), torch.randn((10, Bw, 4), device=dev) * 0.3
    tp = torch.cat([torch.randn((10, Bw, 2), device=dev) * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device=dev)], -1)
    step = lambda: net.step(ob, pp, tp, None, keep_prob=0.5, seed=0)
    for _ in range(3):
        step()
    print("mode %2d  lstm step 2048 windows: %.4f ms" % (mode, timed(step, 10)), flush=True)
    del net
    for Bw2 in (100, 2048):
        n2 = StudentLSTM2(spec=lstm2_spec(units=100, steps=20), seed=1)
        o2, a2 = torch.randn((20, Bw2, 11), device=dev), torch.randn((20, Bw2, 2), device=dev) * 0.3
        t2 = torch.cat([torch.randn((20, Bw2, 2), device=dev) * 0.3, -1 + 0.2 * torch.randn((20, Bw2, 2), device=dev)], -1)
        r2 = torch.randn((20, Bw2), device=dev) * 0.2
        s2 = lambda: n2.step(o2, a2, t2, r2, None, keep_prob=0.5, seed=0)
        for _ in range(3):
            s2()
        print("mode %2d  lstm2 step %d windows: %.4f ms" % (mode, Bw2, timed(s2, 10)), flush=True)
        del n2
    for (M, N, K, a_mn, b_mn, ws) in [(20480, 128, 64, 0, 1, 0), (20480, 200, 64, 0, 0, 0), (244, 800, 20480, 1, 1, 24 * 256 * 800),
                                      (4096, 4096, 4096, 0, 1, 0), (16384, 4096, 4096, 0, 1, 0)]:
        f = gemm(M, N, K, a_mn, b_mn, ws)
        f(); f()
        ms = timed(f, 5)
        print("mode %2d  gemm %6d x %5d x %6d (a_mn %d, b_mn %d): %.4f ms = %.1f TFLOP/s" % (mode, M, N, K, a_mn, b_mn, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
check(lib().rb_gemm_set_cta_packing(-1))
