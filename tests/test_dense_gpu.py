"""Dense stacks on the device (rb_dense_*, rb_vf_targets) vs the float64 restatement: value-function regressor, reward head,
KL-trained stacks of odd widths, ragged batches, Adam loss curve, bit-reproducibility."""
import numpy as np
import pytest
import torch

from oracle import dense_np as DN
from oracle import nn_np as NN

pytestmark = pytest.mark.gpu

KINDS = {"mse": 2, "kl_st": 0, "kl_ts": 1}


def _data(rng, B, dims, kind):
    x = rng.standard_normal((B, dims[0])).astype(np.float32)
    t = (rng.standard_normal((B, dims[-1])) * 0.5).astype(np.float32)
    if kind != "mse":
        t[:, 2:] = -1.0 + 0.2 * t[:, 2:]
    return x, t


@pytest.mark.parametrize("kind,dims,acts,B", [("mse", (13, 64) + (100,) * 10 + (1,), (0,) + (1,) * 10 + (0,), 980),      # the vf scope, 20 episodes x 49
                                              ("mse", (128, 64, 1), (1, 0), 200),                                    # reward head on 128 trunk features
                                              ("kl_st", (11, 33, 17, 4), (1, 1, 0), 1), ("kl_ts", (16, 24, 128, 4), (1, 0, 0), 4097),
                                              ("mse", (7, 5, 3), (1, 0), 129)])
def test_dense_loss_grad_matches_oracle(kind, dims, acts, B):
    from reacherdistilation_b200.vf_train import DenseNet
    rng = np.random.default_rng(B)
    net = DenseNet(dims, acts, seed=2)
    P = net.params.cpu().numpy().astype(np.float64)
    P += rng.standard_normal(P.size) * 0.05                       # non-zero biases
    net.params.copy_(torch.from_numpy(P.astype(np.float32)))
    P = net.params.cpu().numpy().astype(np.float64)
    x, t = _data(rng, B, dims, kind)
    s_ref, l_ref, g_ref = DN.loss_grad(dims, acts, P, x, t, kind)
    fw = net.forward(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(fw - s_ref).max() <= 5e-5 * max(1.0, np.abs(s_ref).max())          # bf16x3 products: ~1e-5 per layer (stated tolerance)
    s = net.loss_grad(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), KINDS[kind]).cpu().numpy()
    gl = net.gradloss.cpu().numpy().astype(np.float64)
    assert np.abs(s - s_ref).max() <= 5e-5 * max(1.0, np.abs(s_ref).max())
    assert abs(gl[-1] - l_ref) <= 2e-4 * max(1.0, abs(l_ref))
    assert np.abs(gl[:-1] - g_ref).max() <= 2e-4 * max(1.0, np.abs(g_ref).max())
    s2 = net.loss_grad(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), KINDS[kind])
    assert torch.equal(net.gradloss, torch.from_numpy(gl.astype(np.float32)).cuda()) and np.array_equal(s2.cpu().numpy(), s)   # bit-reproducible


def test_value_regressor_loss_curve_matches_oracle():
    """minimize_vf_l2 (backup/student_rollout_mlp_vf.py:276,290-295) for 25 steps on add_vtarg targets: device vs float64 restatement + TF Adam."""
    from reacherdistilation_b200.vf_train import ValueRegressor, add_vtarg
    rng = np.random.default_rng(0)
    E, T = 16, 50
    rew = (-np.abs(rng.standard_normal((E, T))) * 0.3).astype(np.float32)
    vt = add_vtarg(torch.from_numpy(rew).cuda(), 0.99).cpu().numpy()
    ref_vt = np.stack([DN.add_vtarg(list(rew[e].astype(np.float64)), 0.99) for e in range(E)])
    assert np.abs(vt - ref_vt).max() <= 2e-6 * max(1.0, np.abs(ref_vt).max())
    prev_ob = rng.standard_normal((E, T - 1, 11)).astype(np.float32)
    next_ac = (rng.standard_normal((E, T - 1, 2)) * 0.3).astype(np.float32)
    targ = vt[:, :T - 1]
    net = ValueRegressor(seed=5, lr=1e-3)
    theta = net.params.cpu().numpy().astype(np.float64)
    opt = NN.AdamTF(theta.size, lr=1e-3)
    x = np.concatenate([prev_ob.reshape(-1, 11), next_ac.reshape(-1, 2)], -1)
    dob, dac, dtg = torch.from_numpy(prev_ob).cuda(), torch.from_numpy(next_ac).cuda(), torch.from_numpy(np.ascontiguousarray(targ)).cuda()
    worst, first, last = 0.0, None, None
    for it in range(25):
        net.fit_step(dob, dac, dtg)
        l_dev = float(net.last_loss())
        _, l_ref, g = DN.loss_grad(net.dims, net.acts, theta.astype(np.float32), x, targ.reshape(-1, 1), "mse")
        theta = opt.update(theta, g)
        worst = max(worst, abs(l_dev - l_ref) / max(1.0, abs(l_ref)))
        first = l_ref if first is None else first
        last = l_ref
    print("value regressor loss curve: first %.4g last %.4g, max rel err %.3g" % (first, last, worst))
    assert last < first and worst <= 2e-3           # stated tolerance: same as the student loss curves


def test_vf_targets_ragged():
    from reacherdistilation_b200.vf_train import add_vtarg
    rng = np.random.default_rng(1)
    for E, T in ((1, 1), (3, 2), (257, 50)):
        r = rng.standard_normal((E, T)).astype(np.float32)
        out = add_vtarg(torch.from_numpy(r).cuda(), 0.9).cpu().numpy()
        ref = np.stack([DN.add_vtarg(list(r[e].astype(np.float64)), 0.9) for e in range(E)])
        assert np.abs(out - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_dense_rejects_bad_specs():
    from reacherdistilation_b200 import ReacherB200Error
    from reacherdistilation_b200.vf_train import DenseNet
    with pytest.raises((ReacherB200Error, ValueError)):
        DenseNet((4, 0, 2))
    net = DenseNet((4, 8, 3), (1, 0))
    x, t = torch.zeros((5, 4), device="cuda"), torch.zeros((5, 3), device="cuda")
    with pytest.raises(ReacherB200Error):
        net.loss_grad(x, t, 0)                      # KL needs a 4-wide output
