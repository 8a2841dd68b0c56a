"""The oracle is pinned to the reference's checked-in TensorBoard graphs (src/~/reacher/data/viz/1/events.out.tfevents.*, written by
lstm_train.py:89-90): tests/golden/make_graph_facts.py extracted the op chain / constants / variable shapes into
tests/golden/graph_facts.json; here oracle/nn_np.py, oracle/lstm_np.py, AdamTF and the product's checkpoint importer are checked
against those facts (CPU only)."""
import json
import os

import numpy as np

from oracle import lstm_np as L
from oracle import nn_np as NN

FACTS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_facts.json")))


def test_facts_come_from_all_twelve_reference_graphs():
    assert len(FACTS["source"]) == 12 and all("events.out.tfevents" in s for s in FACTS["source"])


def test_teacher_graph_matches_oracle_policy():
    t = FACTS["teacher"]
    # layer shapes and activations: 11 -> 64 tanh -> 64 tanh -> 2 linear, logstd [1,2]
    assert [(l["kernel"], l["bias"], l["activation"]) for l in t["layers"]] == [([11, 64], [64], "tanh"), ([64, 64], [64], "tanh"), ([64, 2], [2], "linear")]
    P = NN.policy_unpack(np.zeros(NN.policy_param_count(2), np.float32), 2)
    assert [list(P[k].shape) for k in ("W1", "W2", "W3")] == [l["kernel"] for l in t["layers"]]
    assert P["logstd"].size == int(np.prod(t["logstd_shape"])) and t["obfilter_shapes"]["runningsum"] == [11]
    # clip +-5 on the normalised observation; pdparam = [mean, 0 * mean + logstd]
    assert t["clip"] == [-5.0, 5.0]
    rng = np.random.default_rng(0)
    p = (rng.standard_normal(NN.policy_param_count(2)) * 0.3).astype(np.float32)
    p[:11] = rng.standard_normal(11) * 0.1
    p[11:22] = 0.5 + rng.random(11)
    mu, sd = p[:11].astype(np.float64), p[11:22].astype(np.float64)
    far = NN.policy_fwd((mu + 1e4 * sd)[None], p)
    assert np.allclose(far, NN.policy_fwd((mu + t["clip"][1] * sd)[None], p), atol=1e-12) and not np.allclose(far, NN.policy_fwd((mu + 4.0 * sd)[None], p))
    assert np.allclose(NN.policy_fwd((mu - 1e4 * sd)[None], p), NN.policy_fwd((mu + t["clip"][0] * sd)[None], p), atol=1e-12)
    out = NN.policy_fwd(rng.standard_normal((5, 11)), p)
    assert np.array_equal(out[:, 2:], np.broadcast_to(p[-2:].astype(np.float64), (5, 2)))
    # chain against an independent evaluation in the graph's op order: ((ob - mean) / std) clip -> matmul + bias -> tanh ...
    ob = rng.standard_normal((7, 11))
    z = np.maximum(np.minimum((ob - mu) / sd, t["clip"][1]), t["clip"][0])
    h = z
    P2 = {k: v.astype(np.float64) for k, v in NN.policy_unpack(p, 2).items()}
    for W, b, l in ((P2["W1"], P2["b1"], t["layers"][0]), (P2["W2"], P2["b2"], t["layers"][1]), (P2["W3"], P2["b3"], t["layers"][2])):
        h = h @ W + b
        h = np.tanh(h) if l["activation"] == "tanh" else h
    assert np.allclose(NN.policy_fwd(ob, p)[:, :2], h, atol=1e-14)


def test_obfilter_variance_floor_in_the_checkpoint_importer():
    """std = sqrt(max(E[x^2] - mean^2, floor)) with the graph's floor constant (0.01 as float32)."""
    from reacherdistilation_b200.teacher import policy_params_from_named
    floor = FACTS["teacher"]["obfilter_var_floor"]
    assert abs(floor - 0.01) < 1e-9
    rng = np.random.default_rng(1)
    count, mean = 1000.0, rng.standard_normal(11)
    var = np.concatenate([np.full(5, 1e-4), 0.5 + rng.random(6)])                    # five entries below the floor
    v = {"pi/obfilter/count": count, "pi/obfilter/runningsum": mean * count, "pi/obfilter/runningsumsq": (var + mean ** 2) * count,
         "pi/pol/fc1/kernel": rng.standard_normal((11, 64)), "pi/pol/fc1/bias": np.zeros(64), "pi/pol/fc2/kernel": rng.standard_normal((64, 64)),
         "pi/pol/fc2/bias": np.zeros(64), "pi/pol/final/kernel": rng.standard_normal((64, 2)), "pi/pol/final/bias": np.zeros(2),
         "pi/pol/logstd": np.zeros((1, 2))}
    p = policy_params_from_named(v)
    assert np.allclose(p[11:16], np.sqrt(floor), rtol=1e-6) and np.allclose(p[16:22], np.sqrt(var[5:]), rtol=1e-5)
    assert np.allclose(p[:11], mean, rtol=1e-6)


def test_adam_constants_match_oracle_and_defaults():
    a = FACTS["adam"]
    assert a["use_nesterov"] == [False]
    opt = NN.AdamTF(4)
    assert abs(opt.b1 - a["beta1"]) < 1e-7 and abs(opt.b2 - a["beta2"]) < 1e-7 and abs(opt.eps - a["epsilon"]) < 1e-15
    assert abs(a["learning_rate"] - 1e-3) < 1e-9                                     # the LSTM experiment's rate (lstm_train.py:74)
    import inspect
    from reacherdistilation_b200 import lstm_train, student_nn
    assert inspect.signature(lstm_train.train).parameters["lr"].default == 1e-3
    sig = inspect.signature(student_nn.StudentLSTM.__init__).parameters
    assert (sig["lr"].default, sig["beta1"].default, sig["beta2"].default, sig["eps"].default) == (1e-3, 0.9, 0.999, 1e-8)


def test_lstm_cell_matches_graph_gate_order_forget_bias_and_state_carry():
    f = FACTS["lstm"]
    assert f["gate_order"] == ["i", "j", "f", "o"] and f["forget_bias"] == 1.0 and f["concat_order"] == ["input", "m_prev"]
    assert f["state_carried_through_unroll"] and f["cell_shared_by_steps"] and f["heads_unshared_per_step"]
    rng = np.random.default_rng(2)
    p = L.init_params(3)
    p[L.L_BL:L.L_HEAD0] = rng.standard_normal(L.G).astype(np.float32) * 0.1
    B = 3
    ob, pp = rng.standard_normal((L.T, B, 11)), rng.standard_normal((L.T, B, 4)) * 0.3
    st = rng.standard_normal((2, B, L.U)) * 0.2
    s, fin, cache = L.forward(p, ob, pp, st)
    # independent evaluation of the first two steps straight from the facts
    We, be, Wl, bl, heads = L._views(p)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    c, m = st[0].copy(), st[1].copy()
    for t in range(2):
        x = np.concatenate([ob[t], pp[t] @ We + be], -1)
        z = np.concatenate([x, m], -1) @ Wl + bl                                       # concat_order: input then m_prev, ONE shared kernel
        g = dict(zip(f["gate_order"], np.split(z, 4, axis=1)))
        c = sig(g["f"] + f["forget_bias"]) * c + sig(g["i"]) * np.tanh(g["j"])
        m = sig(g["o"]) * np.tanh(c)
        assert np.allclose(cache[t][6], c, atol=1e-13) and np.allclose(cache[t][7][0], m, atol=1e-13)     # state carried into step t + 1
    # un-shared heads: step 0 and step 1 read different head parameters
    assert not np.array_equal(heads[0][0][0], heads[1][0][0])


def test_kl_is_a_sum_over_every_axis():
    k = FACTS["kl"]
    assert k["reduce"] == "sum" and sorted(k["reduction_indices"]) == [0, 1, 2]
    rng = np.random.default_rng(3)
    s, t = rng.standard_normal((4, 5, 4)) * 0.3, rng.standard_normal((4, 5, 4)) * 0.3
    tot, _ = NN.kl_loss(s, t)
    parts = sum(NN.kl_loss(s[i, j][None], t[i, j][None])[0] for i in range(4) for j in range(5))
    assert abs(tot - parts) < 1e-12 and abs(L.kl(s, t)[0] - tot) < 1e-12


def test_two_headed_graph_topology_of_the_tfevents():
    """Head variables of the recorded two-headed experiment (backup/student_rollout.py:130-170 at an earlier commit): per unrolled step a
    trunk -> 128, a reward head 128 -> 64 -> ... -> 1 and an action head 128 -> 64 -> 4 (pdflat)."""
    hv = FACTS["lstm"]["head_variables"]
    for step in (1, 2):
        assert hv["lstm_step%d/kernel" % step][1] == 128 and hv["reward_hid%d/kernel" % step] == [128, 64]
        assert hv["reward_out%d/kernel" % step] == [64, 1] and hv["lstm_action%d/kernel" % step] == [128, 64] and hv["pd_step%d/kernel" % step] == [64, 4]
    assert FACTS["lstm"]["input_dim"] == 13                                           # dropout(ob) (11) + previous action (2)
