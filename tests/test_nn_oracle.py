"""Host-side checks of the NN oracle: analytic gradients vs finite differences, TF-form Adam vs its closed form."""
import numpy as np

from oracle import nn_np as NN


def test_mlp_gradient_matches_finite_differences():
    rng = np.random.default_rng(0)
    P = (rng.standard_normal(NN.mlp_param_count()) * 0.2).astype(np.float32)
    x, t = rng.standard_normal((9, 16)), rng.standard_normal((9, 4)) * 0.3
    s, hs = NN.mlp_fwd(x, P)
    _, ds = NN.kl_loss(s, t)
    g = NN.mlp_bwd(hs, P, ds)
    for i in rng.integers(0, P.size, 25):
        P1, P2 = P.copy(), P.copy()
        P1[i] += 1e-3; P2[i] -= 1e-3
        fd = (NN.kl_loss(NN.mlp_fwd(x, P1)[0], t)[0] - NN.kl_loss(NN.mlp_fwd(x, P2)[0], t)[0]) / (float(P1[i]) - float(P2[i]))
        assert abs(fd - g[i]) <= 2e-4 * max(1.0, abs(g[i]))


def test_policy_gradient_matches_finite_differences():
    rng = np.random.default_rng(1)
    P = (rng.standard_normal(NN.policy_param_count(4)) * 0.2).astype(np.float32)
    P[:11] = 0; P[11:22] = 1
    x, t = rng.standard_normal((7, 11)), rng.standard_normal((7, 4)) * 0.3
    for loss in (NN.kl_loss, NN.kl_loss_rev):
        s = NN.policy_fwd(x, P, nout=4)
        _, ds = loss(s, t)
        g = NN.policy_bwd(x, P, ds)
        assert (g[:22] == 0).all() and (g[-2:] == 0).all()
        for i in rng.integers(22, P.size - 2, 25):
            P1, P2 = P.copy(), P.copy()
            P1[i] += 1e-3; P2[i] -= 1e-3
            fd = (loss(NN.policy_fwd(x, P1, nout=4), t)[0] - loss(NN.policy_fwd(x, P2, nout=4), t)[0]) / (float(P1[i]) - float(P2[i]))
            assert abs(fd - g[i]) <= 2e-4 * max(1.0, abs(g[i]))


def test_adam_tf_first_step_and_epsilon_placement():
    opt = NN.AdamTF(3, lr=1e-3, eps=1e-8)
    th = opt.update(np.zeros(3), np.array([1.0, -2.0, 0.5]))
    assert np.allclose(th, [-1e-3, 1e-3, -1e-3], rtol=1e-6)       # first step = -lr * sign(g)
    big = NN.AdamTF(1, lr=1e-3, eps=1e-3)                       # MpiAdam epsilon: visible effect on small gradients
    th2 = big.update(np.zeros(1), np.array([1e-3]))
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.isclose(th2[0], -lr_t * 1e-4 / (np.sqrt(0.001 * 1e-6) + 1e-3))


def test_dropout_mask_statistics_and_determinism():
    ids = np.arange(20000, dtype=np.uint32)
    k1 = NN.dropout_keep(5, ids, 3, 0.5)
    k2 = NN.dropout_keep(5, ids, 3, 0.5)
    assert np.array_equal(k1, k2) and set(np.unique(k1)) == {0.0, 1.0}
    assert abs(k1.mean() - 0.5) < 0.01
    assert not np.array_equal(k1, NN.dropout_keep(5, ids, 4, 0.5))
    x = NN.student_input(np.ones((4, 11), np.float32), np.zeros((4, 4)), np.zeros(4), 0.5, 5, ids[:4], 3)
    assert set(np.unique(x[:, :11])) <= {0.0, 2.0} and x.shape == (4, 16)


def test_student_mlp_gradient_matches_torch_autograd_for_every_loss():
    """oracle/nn_np.py (student_nn.py:51-57 graph, loss.py:3-13 KL and the other loss kinds) against torch autograd in float64 on an independently
    written graph: all 24 380 gradients, four loss kinds."""
    import torch
    rng = np.random.default_rng(5)
    P = (rng.standard_normal(NN.mlp_param_count()) * 0.2).astype(np.float32)
    x, t = rng.standard_normal((33, 16)), np.concatenate([rng.standard_normal((33, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((33, 2))], -1)
    dims, tanh = (16, 24, 128, 128, 32, 4), (True, True, False, True, False)        # the third layer is LINEAR (student_nn.py:55)
    for kind in (0, 1, 2, 3):
        s, hs = NN.mlp_fwd(x, P)
        l, ds = NN.pd_loss(s, t, kind)
        g = NN.mlp_bwd(hs, P, ds)
        th = torch.tensor(P.astype(np.float64), requires_grad=True)
        a, o = torch.as_tensor(x), 0
        for i in range(5):
            W = th[o:o + dims[i] * dims[i + 1]].reshape(dims[i], dims[i + 1]); o += dims[i] * dims[i + 1]
            b = th[o:o + dims[i + 1]]; o += dims[i + 1]
            a = a @ W + b
            if tanh[i]:
                a = torch.tanh(a)
        ms, ls, mt, lt = a[:, :2], a[:, 2:], torch.as_tensor(t[:, :2]), torch.as_tensor(t[:, 2:])
        if kind == 0:
            loss = (lt - ls + (torch.exp(2 * ls) + (ms - mt) ** 2) / (2 * torch.exp(2 * lt)) - 0.5).sum()
        elif kind == 1:
            loss = (ls - lt + (torch.exp(2 * lt) + (ms - mt) ** 2) / (2 * torch.exp(2 * ls)) - 0.5).sum()
        elif kind == 2:
            loss = ((a - torch.as_tensor(t)) ** 2).sum()
        else:
            loss = ((ms - mt) ** 2).sum()
        loss.backward()
        assert abs(loss.item() - l) <= 1e-10 * max(1.0, abs(l)), kind
        assert np.abs(th.grad.numpy() - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), kind


def test_policy64_student_gradient_matches_torch_autograd():
    """The 2x64 student of the backup experiment (backup/student_rollout.py:79-87: obfilter clip +-5, two tanh layers, 4 outputs) with both KL
    directions (:639-642): oracle/nn_np.policy_bwd against torch float64 autograd, every trainable parameter; the obfilter and logstd slots of the
    flat layout get no gradient."""
    import torch
    rng = np.random.default_rng(6)
    P = (rng.standard_normal(NN.policy_param_count(4)) * 0.2).astype(np.float32)
    P[:11] = rng.standard_normal(11) * 0.1
    P[11:22] = 0.5 + rng.random(11)
    x = rng.standard_normal((21, 11)) * 3.0                      # some inputs beyond the clip
    t = np.concatenate([rng.standard_normal((21, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((21, 2))], -1)
    for kind, loss_fn in ((0, NN.kl_loss), (1, NN.kl_loss_rev)):
        s = NN.policy_fwd(x, P, nout=4)
        l, ds = loss_fn(s, t)
        g = NN.policy_bwd(x, P, ds)
        th = torch.tensor(P.astype(np.float64), requires_grad=True)
        mu, sd = th[:11], th[11:22]
        o = 22
        W1 = th[o:o + 704].reshape(11, 64); o += 704
        b1 = th[o:o + 64]; o += 64
        W2 = th[o:o + 4096].reshape(64, 64); o += 4096
        b2 = th[o:o + 64]; o += 64
        W3 = th[o:o + 256].reshape(64, 4); o += 256
        b3 = th[o:o + 4]
        z = torch.clamp((torch.as_tensor(x) - mu.detach()) / sd.detach(), -5.0, 5.0)     # the filter statistics are not trained
        a = torch.tanh(torch.tanh(z @ W1 + b1) @ W2 + b2) @ W3 + b3
        ms, ls, mt, lt = a[:, :2], a[:, 2:], torch.as_tensor(t[:, :2]), torch.as_tensor(t[:, 2:])
        loss = (lt - ls + (torch.exp(2 * ls) + (ms - mt) ** 2) / (2 * torch.exp(2 * lt)) - 0.5).sum() if kind == 0 else \
            (ls - lt + (torch.exp(2 * lt) + (ms - mt) ** 2) / (2 * torch.exp(2 * ls)) - 0.5).sum()
        loss.backward()
        ga = th.grad.numpy()
        assert abs(loss.item() - l) <= 1e-10 * max(1.0, abs(l))
        assert np.abs(ga - g).max() <= 1e-9 * max(1.0, np.abs(g).max()) and (g[:22] == 0).all() and (g[-2:] == 0).all()
        assert (np.abs(np.clip((x - P[:11]) / P[11:22], -5, 5)) == 5).any()              # the clip was exercised
