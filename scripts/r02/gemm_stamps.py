"""Phase stamps of CTA 0 of k_gemm_bf16x3 (rb_debug_gemm_stamps) on the LSTM's GEMM shapes, L2-warm (the launch is repeated, last one read)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from reacherdistilation_b200._lib import check, lib, ptr, stream_ptr
dev = "cuda"
L = lib()
names = ["prologue", "k0 copies", "k0 convert", "k1 copies", "k1 convert", "rest of K + MMA", "TMEM->smem", "output", "teardown"]
for (tag, M, N, K, a_mn, b_mn, act, useH, ws) in [("head fwd L2", 20480, 128, 64, 0, 1, 1, 0, 0), ("head fwd L1", 20480, 64, 200, 0, 1, 1, 0, 0),
                                                  ("head fwd L4", 20480, 32, 64, 0, 1, 1, 0, 0), ("head dgrad L1", 20480, 200, 64, 0, 0, 0, 0, 0),
                                                  ("head dgrad L3", 20480, 128, 64, 0, 0, 0, 1, 0), ("W_l gradient", 244, 800, 20480, 1, 1, 0, 0, 24 * 256 * 800)]:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.empty((M, N), device=dev)
    bias = torch.randn(N, device=dev) if act else None
    H = torch.tanh(torch.randn((M, N), device=dev)) if useH else None
    wsb = torch.empty(ws, device=dev) if ws else None
    run = lambda: check(L.rb_gemm_bf16x3(ptr(A), A.shape[1], a_mn, ptr(B), B.shape[1], b_mn, ptr(Cm), N, M, N, K, ptr(bias), act, 0, ptr(H), N if useH else 0,
                                         ptr(wsb), ws, stream_ptr()))
    for mode in (0, 1):
        check(L.rb_gemm_set_cta_packing(mode))
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        buf = (C.c_ulonglong * 16)()
        check(L.rb_debug_gemm_stamps(buf))
        last = (buf[10] - buf[0]) / 1e3, (buf[11] - buf[0]) / 1e3
        st = np.array(list(buf)[:10], dtype=np.int64)
        d = np.diff(st) / 1e3
        print("%-14s %6d x %4d x %6d mode %d: %.1f us per launch, CTA 0 %.1f us: " % (tag, M, N, K, mode, 1e3 * e0.elapsed_time(e1) / 20, (st[9] - st[0]) / 1e3)
              + ", ".join("%s %.2f" % (n, x) for n, x in zip(names, d)) + "; last CTA of the grid: start +%.1f, end +%.1f us" % last, flush=True)
check(L.rb_gemm_set_cta_packing(-1))
