"""Student kernels vs the oracle: forward, fused KL loss + flat gradient, TF-form Adam, dropout mask, determinism."""
import numpy as np
import pytest
import torch

from oracle import nn_np as NN

pytestmark = pytest.mark.gpu


def _modes():
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200._lib import lib
    out = [("fp32", MODE_FP32, 3e-6)]               # measured over all cases below: 8.7e-7 (fp32 kernels) / 2.4e-5 (tcgen05, bf16x3 operands): gates at ~4x
    if lib().rb_student_mode_available(MODE_TC):
        out.append(("tc", MODE_TC, 1e-4))
    return out


def _ref(kind, P, x, t, loss_kind):
    from reacherdistilation_b200 import STUDENT_MLP
    loss = lambda a, b: NN.pd_loss(a, b, loss_kind)     # 0 / 1: the two KL directions, 2: squared error on pdflat, 3: on the actions
    if kind == STUDENT_MLP:
        s, hs = NN.mlp_fwd(x, P)
        l, ds = loss(s, t)
        return s, l, NN.mlp_bwd(hs, P, ds)
    s = NN.policy_fwd(x, P, nout=4)
    l, ds = loss(s, t)
    return s, l, NN.policy_bwd(x, P, ds)


@pytest.mark.parametrize("kind_name", ["mlp", "policy64"])
@pytest.mark.parametrize("B", [1, 63, 64, 65, 200, 1000, 4097])
@pytest.mark.parametrize("loss_kind", [0, 1, 2, 3])
def test_loss_grad_matches_oracle(kind_name, B, loss_kind):
    from reacherdistilation_b200 import STUDENT_MLP, STUDENT_POLICY64
    from reacherdistilation_b200.student_nn import StudentNet
    kind = STUDENT_MLP if kind_name == "mlp" else STUDENT_POLICY64
    rng = np.random.default_rng(B * 7 + loss_kind)
    for name, mode, tol in _modes():
        net = StudentNet(kind=kind, seed=4, mode=mode)
        P = net.params.cpu().numpy().copy()
        if kind == STUDENT_POLICY64:                       # make the 4 outputs non-trivial (normc(0.01) would give ~0)
            P[22 + 704 + 64 + 4096 + 64:-6] *= 30.0
            net.params.copy_(torch.from_numpy(P))
        x = (rng.standard_normal((B, net.in_dim)) * 1.5).astype(np.float32)
        t = np.concatenate([rng.standard_normal((B, 2)) * 0.3, -1.0 + 0.2 * rng.standard_normal((B, 2))], -1).astype(np.float32)
        s_dev = net.loss_grad(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), loss_kind)
        gl = net.gradloss.cpu().numpy().astype(np.float64)
        s, l, g = _ref(kind, P, x, t, loss_kind)
        gscale = max(1.0, np.abs(g).max())
        e_s, e_l, e_g = np.abs(s_dev.cpu().numpy() - s).max(), abs(gl[-1] - l) / max(1.0, abs(l)), np.abs(gl[:-1] - g).max() / gscale
        print("%s %s B=%d loss=%d: s %.3g loss %.3g grad %.3g" % (kind_name, name, B, loss_kind, e_s, e_l, e_g))
        assert e_s <= tol and e_l <= tol and e_g <= tol
        fw = net.forward(torch.from_numpy(x).cuda())
        assert np.abs(fw.cpu().numpy() - s).max() <= tol


def test_gradient_is_bit_reproducible():
    from reacherdistilation_b200 import STUDENT_MLP
    from reacherdistilation_b200.student_nn import StudentNet
    rng = np.random.default_rng(0)
    net = StudentNet(kind=STUDENT_MLP, seed=1)
    x = torch.from_numpy(rng.standard_normal((50000, 16)).astype(np.float32)).cuda()
    t = torch.from_numpy((rng.standard_normal((50000, 4)) * 0.3).astype(np.float32)).cuda()
    net.loss_grad(x, t)
    a = net.gradloss.clone()
    net.loss_grad(x, t)
    assert torch.equal(a, net.gradloss)


@pytest.mark.parametrize("eps,lr", [(1e-8, 1e-4), (1e-3, 1e-3)])
def test_adam_matches_tf_form(eps, lr):
    from reacherdistilation_b200 import STUDENT_MLP
    from reacherdistilation_b200.student_nn import StudentNet
    rng = np.random.default_rng(2)
    net = StudentNet(kind=STUDENT_MLP, seed=1, lr=lr, eps=eps)
    theta = net.params.cpu().numpy().astype(np.float64)
    opt = NN.AdamTF(theta.size, lr=lr, eps=eps)
    for step in range(5):
        g = (rng.standard_normal(theta.size) * 10 ** rng.uniform(-4, 1)).astype(np.float32)
        net.gradloss[:-1].copy_(torch.from_numpy(g))
        net.adam_step(grad_scale=0.5)
        theta = opt.update(theta, g.astype(np.float64) * 0.5)
        assert np.abs(net.params.cpu().numpy() - theta).max() <= 2e-6 * max(1.0, lr / 1e-4)


def test_student_input_and_dropout_mask_bit_exact():
    from reacherdistilation_b200.student_nn import student_mlp_input
    rng = np.random.default_rng(3)
    B, seed, sid0, it = 3000, 77, 123, 9
    obs = rng.standard_normal((B, 11)).astype(np.float32)
    pp = rng.standard_normal((B, 4)).astype(np.float32)
    pr = rng.standard_normal(B).astype(np.float32)
    for kp in (0.5, 0.8, 1.0):
        x = student_mlp_input(torch.from_numpy(obs).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(pr).cuda(), kp, seed, sid0, it)
        ref = NN.student_input(obs, pp, pr, kp, seed, np.arange(B, dtype=np.uint32) + sid0, it, dtype=np.float32)
        assert np.array_equal(x.cpu().numpy(), ref.astype(np.float32)), kp


@pytest.mark.parametrize("kind_name", ["mlp", "policy64"])
def test_fused_step_equals_loss_grad_then_adam(kind_name):
    """rb_student_step (one cooperative launch in RB_MODE_TC) == rb_student_loss_grad followed by rb_adam_step."""
    from reacherdistilation_b200 import STUDENT_MLP, STUDENT_POLICY64
    from reacherdistilation_b200.student_nn import StudentNet
    kind = STUDENT_MLP if kind_name == "mlp" else STUDENT_POLICY64
    rng = np.random.default_rng(5)
    for name, mode, tol in _modes():
        a, b = StudentNet(kind=kind, seed=2, mode=mode, lr=1e-3), StudentNet(kind=kind, seed=2, mode=mode, lr=1e-3)
        for it in range(3):
            B = (300, 5000, 129)[it]
            x = torch.from_numpy((rng.standard_normal((B, a.in_dim))).astype(np.float32)).cuda()
            t = torch.from_numpy(np.concatenate([rng.standard_normal((B, 2)) * 0.3, -1.0 + 0.2 * rng.standard_normal((B, 2))], -1).astype(np.float32)).cuda()
            sa = a.step(x, t, grad_scale=0.5)
            sb = b.loss_grad(x, t)
            b.adam_step(grad_scale=0.5)
            assert torch.equal(sa, sb)
            assert torch.equal(a.gradloss, b.gradloss), name          # same kernel, same order: bit-identical gradient
            assert torch.equal(a.params, b.params), name              # shared adam_update(): bit-identical update
            assert a.t == b.t


def test_tc_gradient_is_bit_reproducible():
    from reacherdistilation_b200 import MODE_TC, STUDENT_MLP
    from reacherdistilation_b200._lib import lib
    from reacherdistilation_b200.student_nn import StudentNet
    if not lib().rb_student_mode_available(MODE_TC):
        pytest.skip("RB_MODE_TC student not built")
    rng = np.random.default_rng(0)
    net = StudentNet(kind=STUDENT_MLP, seed=1, mode=MODE_TC)
    x = torch.from_numpy(rng.standard_normal((50000, 16)).astype(np.float32)).cuda()
    t = torch.from_numpy((rng.standard_normal((50000, 4)) * 0.3).astype(np.float32)).cuda()
    net.loss_grad(x, t)
    a = net.gradloss.clone()
    net.loss_grad(x, t)
    assert torch.equal(a, net.gradloss)
