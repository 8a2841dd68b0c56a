"""Bring-up check of the tcgen05 student kernel: forward, loss, per-block gradient errors vs the float64 restatement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import nn_np as NN
from reacherdistilation_b200 import MODE_FP32, MODE_TC, STUDENT_MLP, STUDENT_POLICY64
from reacherdistilation_b200.student_nn import StudentNet

def ref(kind, P, x, t, lk=0):
    loss = NN.kl_loss if lk == 0 else NN.kl_loss_rev
    if kind == STUDENT_MLP:
        s, hs = NN.mlp_fwd(x, P); l, ds = loss(s, t); return s, l, NN.mlp_bwd(hs, P, ds)
    s = NN.policy_fwd(x, P, nout=4); l, ds = loss(s, t); return s, l, NN.policy_bwd(x, P, ds)

blocks = {STUDENT_MLP: [("W1", 0, 384), ("b1", 384, 408), ("W2", 408, 3480), ("b2", 3480, 3608), ("W3", 3608, 19992), ("b3", 19992, 20120),
                        ("W4", 20120, 24216), ("b4", 24216, 24248), ("W5", 24248, 24376), ("b5", 24376, 24380)],
          STUDENT_POLICY64: [("obf", 0, 22), ("W1", 22, 726), ("b1", 726, 790), ("W2", 790, 4886), ("b2", 4886, 4950), ("W3", 4950, 5206), ("b3", 5206, 5210), ("logstd", 5210, 5212)]}
for kind, name in ((STUDENT_POLICY64, "policy64"), (STUDENT_MLP, "mlp")):
    for B in (128, 200, 1000, 40000):
        rng = np.random.default_rng(B)
        net = StudentNet(kind=kind, seed=4, mode=MODE_TC)
        P = net.params.cpu().numpy().copy()
        if kind == STUDENT_POLICY64:
            P[22 + 704 + 64 + 4096 + 64:-6] *= 30.0
        else:
            P[384:408] = rng.standard_normal(24) * 0.1; P[3480:3608] = rng.standard_normal(128) * 0.1      # non-zero biases
            P[19992:20120] = rng.standard_normal(128) * 0.1; P[24216:24248] = rng.standard_normal(32) * 0.1; P[24376:] = rng.standard_normal(4) * 0.1
        net.params.copy_(torch.from_numpy(P))
        x = (rng.standard_normal((B, net.in_dim)) * 1.5).astype(np.float32)
        t = np.concatenate([rng.standard_normal((B, 2)) * 0.3, -1.0 + 0.2 * rng.standard_normal((B, 2))], -1).astype(np.float32)
        s, l, g = ref(kind, P, x, t)
        fw = net.forward(torch.from_numpy(x).cuda()).cpu().numpy()
        print("%s B=%d fwd err %.3g" % (name, B, np.abs(fw - s).max()), flush=True)
        sd = net.loss_grad(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), 0).cpu().numpy()
        gl = net.gradloss.cpu().numpy().astype(np.float64)
        print("   loss_grad: s err %.3g loss rel %.3g" % (np.abs(sd - s).max(), abs(gl[-1] - l) / max(1, abs(l))), flush=True)
        gs = max(1.0, np.abs(g).max())
        for nm, a, b in blocks[kind]:
            print("      %-6s max|g| %.3g err/gscale %.3g" % (nm, np.abs(g[a:b]).max(), np.abs(gl[a:b] - g[a:b]).max() / gs), flush=True)
