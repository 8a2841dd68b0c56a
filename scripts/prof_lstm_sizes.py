"""LSTM loss_grad (no graph) at a few window-batch sizes -- profiled with an ncu launch list to see how the recurrence kernels scale with
the number of 128-window row blocks (= clusters)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reacherdistilation_b200._lib import lib
from reacherdistilation_b200.student_nn import StudentLSTM
f, b = C.c_int(0), C.c_int(0)
L = lib()
if hasattr(L, "rb_debug_lstm_recur_max_clusters"):
    L.rb_debug_lstm_recur_max_clusters(C.byref(f), C.byref(b))
    print("max active 8-CTA clusters: fwd %d bwd %d" % (f.value, b.value), flush=True)
for Bw in (128, 256, 512, 1024, 2048, 4096):
    net = StudentLSTM(seed=1, max_batch=Bw)
    ob, pp = torch.randn((10, Bw, 11), device="cuda"), torch.randn((10, Bw, 4), device="cuda") * 0.3
    tp = torch.cat([torch.randn((10, Bw, 2), device="cuda") * 0.3, -1 + 0.2 * torch.randn((10, Bw, 2), device="cuda")], -1)
    for _ in range(2):
        net.loss_grad(ob, pp, tp, None, keep_prob=0.5, seed=0, iteration=0)
    torch.cuda.synchronize()
    del net
print("done")
