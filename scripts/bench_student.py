"""Time of the cooperative student kernel (loss_grad) and of the fused step at the config-4 shard and at the full batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reacherdistilation_b200 import MODE_TC, STUDENT_MLP, STUDENT_POLICY64
from reacherdistilation_b200.student_nn import StudentNet
for kind, name in ((STUDENT_MLP, "mlp"), (STUDENT_POLICY64, "policy64")):
    for B in (32768, 262144):
        net = StudentNet(kind=kind, seed=1, mode=MODE_TC)
        x = torch.randn((B, net.in_dim), device="cuda"); t = torch.randn((B, 4), device="cuda") * 0.3
        s = torch.empty((B, 4), device="cuda")
        for fn, nm in ((lambda: net.loss_grad(x, t, s_out=s), "loss_grad"), (lambda: net.step(x, t, s_out=s), "step"), (lambda: net.forward(x, out=s), "forward")):
            for _ in range(5): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): fn()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            print("%-9s B=%6d %-9s %7.1f us  %.3e samples/s" % (name, B, nm, us, B / us * 1e6), flush=True)

# phase breakdown of the cooperative kernel (CTA 0, globaltimer)
import ctypes
from reacherdistilation_b200._lib import lib
names = ["fold+image", "sync", "image load", "tiles", "dump", "sync", "reduce", "sync", "un-fold", "sync+adam", "teardown"]
for B in (128, 32768, 262144):
    net = StudentNet(kind=STUDENT_MLP, seed=1, mode=MODE_TC)
    x = torch.randn((B, 16), device="cuda"); t = torch.randn((B, 4), device="cuda") * 0.3
    for _ in range(5): net.step(x, t)
    buf = (ctypes.c_ulonglong * 48)()
    lib().rb_debug_student_timers(buf)
    ts = [buf[i] for i in range(12)]
    print("B=%d phases (us): " % B + ", ".join("%s %.1f" % (n, (ts[i + 1] - ts[i]) / 1e3) for i, n in enumerate(names)) + "  total %.1f" % ((ts[11] - ts[0]) / 1e3))
    tn = ["x0+sync", "L0 mma", "L0 epi", "L1 mma", "L1 epi", "L2 mma", "L2 epi", "L3 mma", "(skip)", "out epi", "bwd3", "bwd2", "bwd1", "bwd0"]
    tt = [buf[16 + i] for i in range(15)]
    print("   first tile (us): " + ", ".join("%s %.2f" % (n, (tt[i + 1] - tt[i]) / 1e3) for i, n in enumerate(tn) if n != "(skip)" and tt[i + 1] >= tt[i]))
