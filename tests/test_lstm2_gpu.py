"""Two-headed LSTM student (rb_lstm2_*; backup/student_rollout.py:130-200,328) on the device vs the float64 restatement oracle/lstm2_np.py:
forward of both heads, state semantics of the two graph variants, total loss = KL + squared reward error, BPTT gradient of every parameter
block, dropout mask, and a short Adam loss curve."""
import numpy as np
import pytest
import torch

from oracle import lstm2_np as L2
from oracle import nn_np as NN

pytestmark = pytest.mark.gpu
TOL = 3e-5       # bf16x3 tensor-core products vs float64, relative to max(1, |ref|) (same gate as the LSTM student, tests/test_lstm_gpu.py)

SPECS = {
    "source_shipped": L2.SOURCE_SPEC(1, 2),                   # NUM_UNITS 1, STEPS_UNROLLED 2 (:48-50)
    "source_commented": L2.SOURCE_SPEC(100, 20),              # the values commented beside them (:38-41)
    "tfevents": L2.TFEVENTS_SPEC,
    "carried_100x20": (100, 20, 1, 128, 64, 3, 64, 32, 64),
    "odd_widths": (37, 5, 1, 40, 24, 2, 20, 12),
}


def _data(L, B, seed):
    rng = np.random.default_rng(seed)
    ob, ac = rng.standard_normal((L.T, B, 11)).astype(np.float32), (rng.standard_normal((L.T, B, 2)) * 0.3).astype(np.float32)
    tp = np.concatenate([rng.standard_normal((L.T, B, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((L.T, B, 2))], -1).astype(np.float32)
    rt = (rng.standard_normal((L.T, B)) * 0.2).astype(np.float32)
    st = (rng.standard_normal((2, B, L.U)) * 0.3).astype(np.float32)
    return ob, ac, tp, rt, st


def _net(spec, seed):
    from reacherdistilation_b200.student_nn import StudentLSTM2
    L = L2.Layout(spec)
    net = StudentLSTM2(spec=L.spec_array(), seed=seed)
    assert net.P == L.P
    rng = np.random.default_rng(seed)                       # non-zero biases, so every bias path is exercised
    p = net.params.cpu().numpy() + (rng.standard_normal(L.P) * 0.02).astype(np.float32)
    net.params.copy_(torch.from_numpy(p))
    return L, net, p


@pytest.mark.parametrize("name,B", [("source_shipped", 2), ("source_commented", 100), ("tfevents", 2), ("tfevents", 257), ("carried_100x20", 64), ("odd_widths", 9)])
def test_forward_both_heads_and_state(name, B):
    spec = SPECS[name]
    L, net, p = _net(spec, 3)
    ob, ac, _, _, st = _data(L, B, B)
    for state in (None, st):
        s, rew, fin = net.forward(torch.from_numpy(ob).cuda(), torch.from_numpy(ac).cuda(), torch.from_numpy(state).cuda() if state is not None else None)
        rs, rr, rfin, _ = L2.forward(spec, p, ob, ac, state)
        e = max(np.abs(s.cpu().numpy() - rs).max(), np.abs(rew.cpu().numpy() - rr).max(), np.abs(fin.cpu().numpy() - rfin).max())
        print("lstm2 %s B=%d state=%s: forward err %.3g" % (name, B, state is not None, e))
        assert e <= TOL
        if not L.carry and state is not None:
            assert np.array_equal(fin.cpu().numpy(), st)       # final_state is the fed state in the source's graph (:156,191)


@pytest.mark.parametrize("name,B,keep_prob", [("source_shipped", 2, 1.0), ("source_commented", 100, 0.5), ("tfevents", 2, 1.0), ("tfevents", 300, 0.8),
                                              ("carried_100x20", 100, 0.5), ("odd_widths", 9, 1.0)])
def test_total_loss_and_bptt_gradient(name, B, keep_prob):
    spec = SPECS[name]
    L, net, p = _net(spec, 4)
    ob, ac, tp, rt, st = _data(L, B, 10 * B + 1)
    seed, sid0, it = 9, 1000, 3
    s, rew = net.loss_grad(torch.from_numpy(ob).cuda(), torch.from_numpy(ac).cuda(), torch.from_numpy(tp).cuda(), torch.from_numpy(rt).cuda(),
                           torch.from_numpy(st).cuda(), keep_prob=keep_prob, seed=seed, sample_id0=sid0, iteration=it)
    gl = net.gradloss.cpu().numpy().astype(np.float64)
    ids = np.arange(L.T * B, dtype=np.uint32) + sid0
    obd = NN.student_input(ob.reshape(-1, 11), np.zeros((L.T * B, 4)), np.zeros(L.T * B), keep_prob, seed, ids, it, dtype=np.float32)[:, :11]
    obd = obd.reshape(L.T, B, 11)
    if keep_prob < 1.0:
        assert (obd == 0).mean() > 0.05                       # the mask is really applied
    rs, rr, (tot, lk, lr), rg = L2.loss_grad(spec, p, obd, ac, tp, rt, st)
    e_s, e_r = np.abs(s.cpu().numpy() - rs).max(), np.abs(rew.cpu().numpy() - rr).max()
    e_l = max(abs(gl[L.P] - tot) / max(1.0, abs(tot)), abs(gl[L.P + 1] - lk) / max(1.0, abs(lk)), abs(gl[L.P + 2] - lr) / max(1.0, abs(lr)))
    e_g = np.abs(gl[:L.P] - rg).max() / max(1.0, np.abs(rg).max())
    print("lstm2 %s B=%d kp=%.1f: s %.3g reward %.3g loss %.3g grad %.3g (|g|max %.3g)" % (name, B, keep_prob, e_s, e_r, e_l, e_g, np.abs(rg).max()))
    assert e_s <= TOL and e_r <= TOL and e_l <= TOL and e_g <= TOL
    # every parameter block against ITS OWN largest gradient entry (cell kernel / bias, and each head layer of the first and last step)
    blocks = [("W_l", L.o_Wl, L.o_bl), ("b_l", L.o_bl, L.head0)]
    for t in (0, L.T - 1):
        base = L.head0 + t * L.head_sz
        for nm, ow, ob_, fi, fo in L.blocks:
            blocks += [("%s%d/W" % (nm, t), base + ow, base + ob_), ("%s%d/b" % (nm, t), base + ob_, base + ob_ + fo)]
    worst = 0.0
    for nm, a, b in blocks:
        scale = np.abs(rg[a:b]).max()
        if nm.startswith("reward_out") and nm.endswith("/b"):      # ONE entry = sum_b 2 (reward - target): judged against the sum of the terms' magnitudes
            scale = max(scale, (2 * np.abs(rr - rt)[int(nm[len("reward_out"):-2])]).sum())
        e_b = np.abs(gl[a:b] - rg[a:b]).max() / max(1e-3, scale)
        worst = max(worst, e_b)
        assert e_b <= 10 * TOL, nm
    print("  worst per-block gradient error (relative to the block's own max) %.3g" % worst)


def test_source_variant_steps_do_not_see_each_other():
    L, net, p = _net(SPECS["source_commented"], 5)
    ob, ac, _, _, st = _data(L, 8, 0)
    f = lambda o: net.forward(torch.from_numpy(o).cuda(), torch.from_numpy(ac).cuda(), torch.from_numpy(st).cuda())
    s, r, _ = f(ob)
    ob2 = ob.copy(); ob2[0] += 1.0
    s2, r2, _ = f(ob2)
    assert torch.equal(s[1:], s2[1:]) and torch.equal(r[1:], r2[1:]) and not torch.equal(s[0], s2[0])


def test_adam_loss_curve_matches_oracle():
    """12 optimiser steps (loss_grad + TF-form Adam, lr 1e-3: :331-336) on fresh windows each step vs the float64 chain."""
    spec = SPECS["carried_100x20"]
    L, net, p = _net(spec, 6)
    theta = p.astype(np.float64)
    opt = NN.AdamTF(L.P, lr=1e-3, eps=1e-8)
    B, worst = 32, 0.0
    first = last = None
    for it in range(12):
        ob, ac, tp, rt, _ = _data(L, B, 100 + it)
        net.loss_grad(torch.from_numpy(ob).cuda(), torch.from_numpy(ac).cuda(), torch.from_numpy(tp).cuda(), torch.from_numpy(rt).cuda(), None,
                      keep_prob=0.5, seed=2, sample_id0=0, iteration=it)
        dev = float(net.gradloss[L.P])
        net.adam_step()
        ids = np.arange(L.T * B, dtype=np.uint32)
        obd = NN.student_input(ob.reshape(-1, 11), np.zeros((L.T * B, 4)), np.zeros(L.T * B), 0.5, 2, ids, it, dtype=np.float32)[:, :11].reshape(L.T, B, 11)
        _, _, (tot, _, _), g = L2.loss_grad(spec, theta.astype(np.float32), obd, ac, tp, rt, None)
        theta = opt.update(theta, g)
        worst = max(worst, abs(dev - tot) / max(1.0, abs(tot)))
        first, last = (tot if first is None else first), tot
    print("lstm2 loss curve: first %.5g last %.5g, max rel err %.3g" % (first, last, worst))
    assert worst <= 1e-4 and last < first


def test_lstm_train_loop_learns_saves_and_restores(tmp_path):
    """backup/student_rollout.py:260-588 as lstm2_train.lstm_train: teacher warm-up, one optimiser step per env step on random windows, the
    student acting from its carried state; checkpoint every episode; -r continues from it; replay mode restores the student only."""
    from reacherdistilation_b200 import lstm2_train
    ck = str(tmp_path / "lstm2.pt")
    kw = dict(num_envs=16, batch_size=32, units=16, steps=5, carry_state=True, verbose=False, seed=3)
    out = lstm2_train.lstm_train(True, 0.5, ck, False, total_episodes=3, **kw)
    assert out["episodes"] >= 4 and len(out["losses"]) == len(out["rets"]) == out["episodes"] - 1 and np.isfinite(out["losses"]).all()
    assert out["losses"][-1] < out["losses"][0] and out["student"].t == out["iterations"]
    p_end, t_end = out["student"].params.clone(), out["student"].t
    out["env"].close()
    again = lstm2_train.lstm_train(True, 0.5, ck, True, iterations=10, **kw)
    assert again["student"].t == t_end + 10 and again["episodes"] == out["episodes"]         # history + optimiser state came back
    again["env"].close()
    replay = lstm2_train.lstm_train(False, 1.0, ck, True, **kw)
    assert replay["student"].t == t_end + 10 and not torch.equal(replay["student"].params, p_end)
    replay["env"].close()
    src = lstm2_train.lstm_train(True, 1.0, str(tmp_path / "src.pt"), False, num_envs=4, batch_size=2, units=1, steps=2, iterations=20, verbose=False)
    assert np.isfinite(src["last_loss"]) and src["student"].P == 34246                         # the source's shipped sizes (:48-50)
    src["env"].close()


def test_graph_step_equals_loss_grad_plus_adam():
    """rb_lstm2_step (one CUDA-graph launch, device-side clock for the dropout iteration and the Adam step) vs rb_lstm2_loss_grad + rb_adam_step:
    bit-identical parameters and losses over 6 steps on changing data in static buffers."""
    from reacherdistilation_b200.student_nn import StudentLSTM2
    spec = SPECS["source_commented"]
    L = L2.Layout(spec)
    a, b = StudentLSTM2(spec=L.spec_array(), seed=7), StudentLSTM2(spec=L.spec_array(), seed=7)
    B = 100
    bufs = [torch.empty((L.T, B, 11), device="cuda"), torch.empty((L.T, B, 2), device="cuda"), torch.empty((L.T, B, 4), device="cuda"), torch.empty((L.T, B), device="cuda")]
    for it in range(6):
        ob, ac, tp, rt, _ = _data(L, B, 50 + it)
        for dst, src in zip(bufs, (ob, ac, tp, rt)):
            dst.copy_(torch.from_numpy(src))
        a.loss_grad(bufs[0], bufs[1], bufs[2], bufs[3], None, keep_prob=0.5, seed=3, sample_id0=0, iteration=a.t)
        la = float(a.gradloss[L.P])
        a.adam_step()
        b.step(bufs[0], bufs[1], bufs[2], bufs[3], None, keep_prob=0.5, seed=3, sample_id0=0)
        assert float(b.gradloss[L.P]) == la, it
        assert torch.equal(a.params, b.params) and torch.equal(a.m, b.m) and torch.equal(a.v, b.v), it
    assert a.t == b.t == 6
