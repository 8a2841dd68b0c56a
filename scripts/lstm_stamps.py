"""Phase stamps of CTA 0 of the LSTM recurrence kernels (rb_debug_lstm_recur_stamps), averaged over the interior steps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reacherdistilation_b200._lib import lib
from reacherdistilation_b200.student_nn import StudentLSTM
L = lib()
for Bw in (128, 1920, 2048):
    net = StudentLSTM(seed=1, max_batch=Bw)
    ob, pp = torch.randn((10, Bw, 11), device="cuda"), torch.randn((10, Bw, 4), device="cuda") * 0.3
    tp = torch.cat([torch.randn((10, Bw, 2), device="cuda") * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device="cuda")], -1)
    for _ in range(3):
        net.loss_grad(ob, pp, tp, None, keep_prob=0.5, seed=0, iteration=0)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 320)()
    L.rb_debug_lstm_recur_stamps(buf)
    st = np.array(list(buf), dtype=np.int64).reshape(2, 10, 16)
    f, b = st[0], st[1]
    print("B=%d forward: total %.1f us" % (Bw, (f[9, 7] - f[0, 0]) / 1e3))
    d = np.diff(f[1:9, :9], axis=1).mean(0) / 1e3
    print("  per step (us): q0 %.2f q1 %.2f q2 %.2f q3 %.2f mma-done %.2f cell %.2f store %.2f cluster %.2f | step %.2f"
          % (*d, (f[2:9, 0] - f[1:8, 0]).mean() / 1e3))
    print("B=%d backward: total %.1f us" % (Bw, (b[0, 7] - b[9, 0]) / 1e3))
    d = np.diff(b[1:9, :8], axis=1).mean(0) / 1e3
    print("  per step (us): loads %.2f dz+A %.2f stage %.2f dz-store %.2f mma-wait %.2f partial-store %.2f cluster %.2f | step %.2f"
          % (*d, (b[1:8, 0] - b[2:9, 0]).mean() / 1e3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net.step(ob, pp, tp, None, keep_prob=0.5, seed=0)
    e1.record(); torch.cuda.synchronize()
    print("  optimiser step (graph): %.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)
    del net
