"""Phase stamps (CTA 0, globaltimer) of k_student_tc at the config-4 shard: launch phases and the phases of the first tile."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from reacherdistilation_b200 import MODE_TC, STUDENT_MLP, _lib
from reacherdistilation_b200.mlp_train import DaggerTrainer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
tr = DaggerTrainer(num_envs=n, seed=0, student_kind=STUDENT_MLP, mode=MODE_TC, keep_prob=kp)
for _ in range(10):
    tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    tr.step()
e1.record(); torch.cuda.synchronize()
print("dagger iteration (keep_prob %.1f, %d envs): %.2f us" % (kp, n, e0.elapsed_time(e1) * 1e3 / 50))
tb = (ctypes.c_ulonglong * 48)()
_lib.lib().rb_debug_student_timers(tb)
us = lambda i, j: (tb[j] - tb[i]) / 1e3
print("launch: setup %.2f image %.2f tiles %.2f dump+act %.2f sync1 %.2f reduce %.2f sync2 %.2f unfold+adam %.2f teardown %.2f total %.2f"
      % (us(0, 1), us(1, 3), us(3, 4), us(4, 5), us(5, 6), us(6, 7), us(7, 8), us(8, 9), us(9, 11), us(0, 11)))
names = ["X0", "fwd0 mma", "epi0", "fwd1 mma", "epi1", "fwd2 mma", "epi2", "fwd3 mma", "out epi", "bwd3", "bwd2", "bwd1", "bwd0"]
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14]
print("first tile (last pass):", " | ".join("%s %.2f" % (names[k], us(16 + order[k], 16 + order[k + 1])) for k in range(len(order) - 1) if tb[16 + order[k + 1]] > tb[16 + order[k]]))
lg = lambda: tr.student.loss_grad(tr.x, tr.t_pd, s_out=tr.s_train if tr.s_train is not None else tr.s_pd)
for _ in range(3): lg()
torch.cuda.synchronize(); e0.record()
for _ in range(20): lg()
e1.record(); torch.cuda.synchronize()
print("loss_grad alone (image kernel + cooperative kernel): %.2f us" % (e0.elapsed_time(e1) * 1e3 / 20))
