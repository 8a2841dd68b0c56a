"""The float64 restatement of the LSTM student (oracle/lstm_np.py): its hand-derived BPTT against finite differences."""
import numpy as np

from oracle import lstm_np as L


def test_bptt_matches_finite_differences():
    rng = np.random.default_rng(0)
    p = L.init_params(1).astype(np.float64)
    p += rng.standard_normal(p.size) * 0.01
    B = 3
    obd, pp = rng.standard_normal((L.T, B, 11)), rng.standard_normal((L.T, B, 4)) * 0.3
    tp = np.concatenate([rng.standard_normal((L.T, B, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((L.T, B, 2))], -1)
    st = rng.standard_normal((2, B, 200)) * 0.1
    for rev in (False, True):
        s, l, g = L.loss_grad(p, obd, pp, tp, st, reverse=rev)
        idx = np.concatenate([rng.choice(p.size, 5, replace=False), [5, 130, 200, L.L_BL + 3, L.L_HEAD0 + 10, L.L_HEAD0 + 9 * L.L_HEAD_SZ + 31650]])
        for k in idx:
            e, q = 1e-6, p.copy()
            q[k] += e
            lp = L.kl(L.forward(q, obd, pp, st)[0], tp, rev)[0]
            q[k] -= 2 * e
            lm = L.kl(L.forward(q, obd, pp, st)[0], tp, rev)[0]
            assert abs((lp - lm) / (2 * e) - g[k]) <= 1e-5 * max(1.0, abs(g[k]))
    assert L.param_count() == 511880


def test_bptt_matches_torch_autograd_on_every_parameter():
    """oracle/lstm_np.py's hand-derived BPTT against torch autograd (float64) on the same graph written independently: all 511 880 gradients."""
    import torch
    rng = np.random.default_rng(4)
    p = L.init_params(2).astype(np.float64) + rng.standard_normal(L.P) * 0.01
    B = 4
    obd, pp = rng.standard_normal((L.T, B, 11)), rng.standard_normal((L.T, B, 4)) * 0.3
    tp = np.concatenate([rng.standard_normal((L.T, B, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((L.T, B, 2))], -1)
    st = rng.standard_normal((2, B, 200)) * 0.1
    for rev in (False, True):
        _, loss, g = L.loss_grad(p, obd, pp, tp, st, reverse=rev)
        th = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        We, be = th[L.L_WE:L.L_BE].reshape(4, L.E), th[L.L_BE:L.L_WL]
        Wl, bl = th[L.L_WL:L.L_BL].reshape(L.XH, L.G), th[L.L_BL:L.L_HEAD0]
        c, m = torch.as_tensor(st[0]), torch.as_tensor(st[1])
        total = 0.0
        for t in range(L.T):
            x = torch.cat([torch.as_tensor(obd[t]), torch.as_tensor(pp[t]) @ We + be, m], -1)
            z = x @ Wl + bl
            U = L.U
            i, j, f, o = torch.sigmoid(z[:, :U]), torch.tanh(z[:, U:2 * U]), torch.sigmoid(z[:, 2 * U:3 * U] + 1.0), torch.sigmoid(z[:, 3 * U:])
            c = f * c + i * j
            m = o * torch.tanh(c)
            a, off = m, L.L_HEAD0 + t * L.L_HEAD_SZ
            for l in range(5):
                W = th[off:off + L.HD[l] * L.HD[l + 1]].reshape(L.HD[l], L.HD[l + 1]); off += L.HD[l] * L.HD[l + 1]
                b = th[off:off + L.HD[l + 1]]; off += L.HD[l + 1]
                a = a @ W + b
                if l < 4:
                    a = torch.tanh(a)
            ms, ls = a[:, :2], a[:, 2:]
            mt, lt = torch.as_tensor(tp[t][:, :2]), torch.as_tensor(tp[t][:, 2:])
            if not rev:
                total = total + (lt - ls + (torch.exp(2 * ls) + (ms - mt) ** 2) / (2 * torch.exp(2 * lt)) - 0.5).sum()
            else:
                total = total + (ls - lt + (torch.exp(2 * lt) + (ms - mt) ** 2) / (2 * torch.exp(2 * ls)) - 0.5).sum()
        total.backward()
        ga = th.grad.numpy()
        assert abs(total.item() - loss) <= 1e-10 * max(1.0, abs(loss))
        assert np.abs(ga - g).max() <= 1e-9 * max(1.0, np.abs(g).max())
