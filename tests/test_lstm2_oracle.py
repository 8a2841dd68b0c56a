"""The float64 restatement of the two-headed LSTM student (oracle/lstm2_np.py; backup/student_rollout.py:130-200,303-328): its layout against the
variable shapes of the reference's checked-in TensorBoard graphs (tests/golden/graph_facts.json), the two graph variants' state semantics, and
its hand-derived BPTT against finite differences.  CPU only."""
import json
import os

import numpy as np

from oracle import lstm2_np as L2

FACTS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_facts.json")))


def _data(L, B, seed):
    rng = np.random.default_rng(seed)
    ob, ac = rng.standard_normal((L.T, B, 11)), rng.standard_normal((L.T, B, 2)) * 0.3
    tp = np.concatenate([rng.standard_normal((L.T, B, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((L.T, B, 2))], -1)
    rt = rng.standard_normal((L.T, B)) * 0.2
    st = rng.standard_normal((2, B, L.U)) * 0.3
    return ob, ac, tp, rt, st


def test_layout_matches_the_recorded_tfevents_graph():
    f = FACTS["lstm"]
    L = L2.Layout(L2.TFEVENTS_SPEC)
    assert f["kernel_shape"] == [L.XH, L.G] and f["units"] == L.U and f["unrolled_steps"] == L.T and f["input_dim"] == L2.IN
    assert f["state_carried_through_unroll"] is True and L.carry == 1 and f["heads_unshared_per_step"] is True
    want = {"trunk": "lstm_step", "reward0": "reward_hid", "reward1": "reward_2hid", "reward2": "reward_3hid", "reward_out": "reward_out",
            "action": "lstm_action", "pd": "pd_step"}
    hv = f["head_variables"]
    for step in (1, 2):
        for name, _, _, fi, fo in L.blocks:
            assert hv["%s%d/kernel" % (want[name], step)] == [fi, fo] and hv["%s%d/bias" % (want[name], step)] == [fo]
    assert len(hv) == 2 * 2 * len(L.blocks)                                    # nothing in the recorded heads is left unaccounted for
    assert L.P == L.XH * L.G + L.G + L.T * sum(fi * fo + fo for _, _, _, fi, fo in L.blocks)


def test_source_variant_never_carries_the_state():
    """backup/student_rollout.py:156 `output, next_state = cell(x_i, state)`: `state` is not reassigned, so step tau's output only depends on the
    window's row tau and the fed state, and final_state (:191) is the fed state."""
    spec = L2.SOURCE_SPEC(units=5, steps=3)
    L = L2.Layout(spec)
    p = L2.init_params(spec, 2)
    ob, ac, _, _, st = _data(L, 4, 0)
    s, r, fin, _ = L2.forward(spec, p, ob, ac, st)
    assert np.array_equal(fin, st)
    ob2 = ob.copy(); ob2[0] += 1.0                                              # disturbing row 0 leaves the later steps untouched
    s2, r2, _, _ = L2.forward(spec, p, ob2, ac, st)
    assert np.array_equal(s2[1:], s[1:]) and np.array_equal(r2[1:], r[1:]) and not np.array_equal(s2[0], s[0])
    carried = (5, 3, 1) + spec[3:]
    s3, _, fin3, _ = L2.forward(carried, p, ob2, ac, st)
    s4, _, _, _ = L2.forward(carried, p, ob, ac, st)
    assert not np.array_equal(s3[1:], s4[1:]) and not np.array_equal(fin3, st)
    assert np.array_equal(s3[0], s2[0])                                          # the first step is the same graph in both variants


def test_bptt_matches_finite_differences():
    for spec in (L2.SOURCE_SPEC(units=6, steps=3), (6, 3, 1, 128, 64, 1, 64), (3, 4, 1, 16, 8, 3, 8, 4, 8), L2.TFEVENTS_SPEC):
        L = L2.Layout(spec)
        rng = np.random.default_rng(1)
        p = L2.init_params(spec, 1).astype(np.float64)
        p += rng.standard_normal(p.size) * 0.02
        ob, ac, tp, rt, st = _data(L, 3, 5)
        s, rew, (tot, lk, lr), g = L2.loss_grad(spec, p, ob, ac, tp, rt, st)
        assert abs(tot - (lk + lr)) < 1e-12 and lk > 0 and lr > 0
        def total(q):
            s_, r_, _, _ = L2.forward(spec, q, ob, ac, st)
            return L2.kl(s_, tp)[0] + ((r_ - rt) ** 2).sum()
        named = [L.o_bl + 1, L.head0 + 3] + [L.head0 + (L.T - 1) * L.head_sz + ow + 1 for _, ow, _, _, _ in L.blocks] \
            + [L.head0 + ob_ for _, _, ob_, _, _ in L.blocks]
        idx = np.concatenate([rng.choice(L.o_bl, 6, replace=False), named])
        for k in idx:
            e, q = 1e-6, p.copy()
            q[k] += e
            lp = total(q)
            q[k] -= 2 * e
            lm = total(q)
            assert abs((lp - lm) / (2 * e) - g[k]) <= 2e-6 * max(1.0, abs(g[k])), (spec, k)


def _torch_total_loss(L, theta, ob, ac, tp, rt, st):
    """The same graph written independently with torch float64 ops (no code shared with oracle/lstm2_np.py beyond the layout): autograd gives
    the gradient of EVERY parameter, against which the oracle's hand-derived back-propagation is checked."""
    import torch
    U, T = L.U, L.T
    Wl, bl = theta[L.o_Wl:L.o_bl].reshape(L.XH, L.G), theta[L.o_bl:L.head0]
    c0, m0 = torch.as_tensor(st[0]), torch.as_tensor(st[1])
    c, m = c0, m0
    total = 0.0
    for t in range(T):
        c_prev, m_prev = (c, m) if L.carry else (c0, m0)
        x = torch.cat([torch.as_tensor(ob[t]), torch.as_tensor(ac[t]), m_prev], -1)
        z = x @ Wl + bl
        i, j, f, o = torch.sigmoid(z[:, :U]), torch.tanh(z[:, U:2 * U]), torch.sigmoid(z[:, 2 * U:3 * U] + 1.0), torch.sigmoid(z[:, 3 * U:])
        c = f * c_prev + i * j
        m = o * torch.tanh(c)
        base = L.head0 + t * L.head_sz
        lay = {name: (theta[base + ow:base + ow + fi * fo].reshape(fi, fo), theta[base + ob_:base + ob_ + fo]) for name, ow, ob_, fi, fo in L.blocks}
        trunk = torch.tanh(m @ lay["trunk"][0] + lay["trunk"][1])
        a = trunk
        for k in range(L.nR):
            a = torch.tanh(a @ lay["reward%d" % k][0] + lay["reward%d" % k][1])
        rew = (a @ lay["reward_out"][0] + lay["reward_out"][1])[:, 0]
        s = torch.tanh(trunk @ lay["action"][0] + lay["action"][1]) @ lay["pd"][0] + lay["pd"][1]
        ms, ls = s[:, :2], s[:, 2:]
        mt, lt = torch.as_tensor(tp[t][:, :2]), torch.as_tensor(tp[t][:, 2:])
        kl = (lt - ls + (torch.exp(2 * ls) + (ms - mt) ** 2) / (2 * torch.exp(2 * lt)) - 0.5).sum()      # lstm_loss, backup/student_rollout.py:196-200
        total = total + kl + ((rew - torch.as_tensor(rt[t])) ** 2).sum()                                 # + reward term, :328
    return total


def test_bptt_matches_torch_autograd_on_every_parameter():
    import torch
    for spec in (L2.SOURCE_SPEC(units=7, steps=4), (7, 4, 1, 128, 64, 1, 64), L2.TFEVENTS_SPEC, (5, 3, 1, 24, 16, 3, 16, 8, 16)):
        L = L2.Layout(spec)
        rng = np.random.default_rng(3)
        p = L2.init_params(spec, 2).astype(np.float64) + rng.standard_normal(L.P) * 0.02
        ob, ac, tp, rt, st = _data(L, 5, 9)
        _, _, (tot, _, _), g = L2.loss_grad(spec, p, ob, ac, tp, rt, st)
        theta = torch.tensor(p, dtype=torch.float64, requires_grad=True)
        loss = _torch_total_loss(L, theta, ob, ac, tp, rt, st)
        loss.backward()
        ga = theta.grad.numpy()
        assert abs(loss.item() - tot) <= 1e-10 * max(1.0, abs(tot))
        assert np.abs(ga - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), spec
        assert (np.abs(g) > 0).mean() > 0.5                                 # the comparison is not vacuous
