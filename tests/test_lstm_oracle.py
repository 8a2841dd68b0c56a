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
