"""The dense-stack oracle (oracle/dense_np.py): analytic gradients vs central finite differences; add_vtarg known answers."""
import numpy as np
import pytest

from oracle import dense_np as DN


@pytest.mark.parametrize("kind,dims,acts", [("mse", (13, 64, 20, 20, 1), (0, 1, 1, 0)), ("kl_st", (11, 16, 8, 4), (1, 1, 0)),
                                            ("kl_ts", (5, 7, 4), (1, 0)), ("mse", (6, 9, 3), (1, 0))])
def test_gradient_matches_finite_differences(kind, dims, acts):
    rng = np.random.default_rng(3)
    n = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
    P = rng.standard_normal(n) * 0.4
    x = rng.standard_normal((9, dims[0]))
    t = rng.standard_normal((9, dims[-1])) * 0.5
    if kind != "mse":
        t[:, 2:] = -1.0 + 0.2 * t[:, 2:]
    _, l0, g = DN.loss_grad(dims, acts, P, x, t, kind)
    idx = rng.choice(n, size=40, replace=False)
    for i in idx:
        e = np.zeros(n); e[i] = 1e-6
        lp = DN.loss_grad(dims, acts, P + e, x, t, kind)[1]
        lm = DN.loss_grad(dims, acts, P - e, x, t, kind)[1]
        fd = (lp - lm) / 2e-6
        assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_add_vtarg_known_answers():
    # T = 3, gamma = 0.5: target[2] = 0.5^3 r2 ; target[1] = 0.5^1 r1 + target[2] ; target[0] = r0 + target[1]   (backup/student_rollout_mlp_vf.py:608-616)
    r = [1.0, 2.0, 4.0]
    assert np.allclose(DN.add_vtarg(r, 0.5), [1.0 + 1.0 + 0.5, 1.0 + 0.5, 0.5])
    r = np.arange(50, dtype=np.float64) - 20.0
    tgt = DN.add_vtarg(list(r), 0.99)
    assert tgt[-1] == pytest.approx(0.99 ** 50 * r[-1])
    assert tgt[0] == pytest.approx(sum(0.99 ** i * r[i] for i in range(49)) + 0.99 ** 50 * r[49])
