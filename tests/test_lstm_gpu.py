"""LSTM student on the device (tcgen05 GEMMs + element-wise kernels) vs the float64 restatement: forward, carried state, KL loss,
BPTT gradient of all 511 880 parameters, dropout mask, and the training loop entry point."""
import numpy as np
import pytest
import torch

from oracle import lstm_np as L
from oracle import nn_np as NN

pytestmark = pytest.mark.gpu
TOL = 3e-5      # tensor-core (bf16x3) LSTM kernels vs float64, relative to max(1, |ref|): measured 7.7e-6 over the cases below, gate at ~4x


def _data(B, seed):
    rng = np.random.default_rng(seed)
    ob = rng.standard_normal((L.T, B, 11)).astype(np.float32)
    pp = (rng.standard_normal((L.T, B, 4)) * 0.3).astype(np.float32)
    tp = np.concatenate([rng.standard_normal((L.T, B, 2)) * 0.3, -1 + 0.2 * rng.standard_normal((L.T, B, 2))], -1).astype(np.float32)
    st = (rng.standard_normal((2, B, 200)) * 0.1).astype(np.float32)
    return ob, pp, tp, st


@pytest.mark.parametrize("B", [1, 20, 150])
def test_forward_and_state_match_oracle(B):
    from reacherdistilation_b200.student_nn import StudentLSTM
    net = StudentLSTM(seed=3)
    p = net.params.cpu().numpy()
    ob, pp, tp, st = _data(B, B)
    for state in (None, st):
        s, fin = net.forward(torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(state).cuda() if state is not None else None)
        rs, rfin, _ = L.forward(p, ob, pp, state)
        assert np.abs(s.cpu().numpy() - rs).max() <= 1e-4 and np.abs(fin.cpu().numpy() - rfin).max() <= 1e-4


@pytest.mark.parametrize("B,keep_prob,loss_kind", [(20, 1.0, 0), (20, 0.5, 0), (7, 1.0, 1), (300, 0.8, 0)])
def test_loss_and_bptt_gradient_match_oracle(B, keep_prob, loss_kind):
    from reacherdistilation_b200.student_nn import StudentLSTM
    net = StudentLSTM(seed=4)
    p = net.params.cpu().numpy()
    ob, pp, tp, st = _data(B, 10 * B)
    seed, sid0, it = 9, 1000, 3
    s = net.loss_grad(torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(tp).cuda(), torch.from_numpy(st).cuda(),
                      keep_prob=keep_prob, seed=seed, sample_id0=sid0, iteration=it, loss_kind=loss_kind)
    gl = net.gradloss.cpu().numpy().astype(np.float64)
    ids = np.arange(L.T * B, dtype=np.uint32) + sid0
    obd = NN.student_input(ob.reshape(-1, 11), np.zeros((L.T * B, 4)), np.zeros(L.T * B), keep_prob, seed, ids, it, dtype=np.float32)[:, :11]
    rs, rl, rg = L.loss_grad(p, obd.reshape(L.T, B, 11), pp, tp, st, reverse=bool(loss_kind))
    gs = max(1.0, np.abs(rg).max())
    e_s, e_l, e_g = np.abs(s.cpu().numpy() - rs).max(), abs(gl[-1] - rl) / max(1.0, abs(rl)), np.abs(gl[:-1] - rg).max() / gs
    print("lstm B=%d kp=%.1f loss=%d: s %.3g loss %.3g grad %.3g (|g|max %.3g)" % (B, keep_prob, loss_kind, e_s, e_l, e_g, gs))
    assert e_s <= TOL and e_l <= TOL and e_g <= TOL
    blocks = dict(We=(0, 128), be=(128, 160), Wl=(160, L.L_BL), bl=(L.L_BL, L.L_HEAD0), head0=(L.L_HEAD0, L.L_HEAD0 + L.L_HEAD_SZ),
                  head9=(L.L_HEAD0 + 9 * L.L_HEAD_SZ, L.P))
    worst_block = 0.0
    for name, (a, b) in blocks.items():
        e_b = np.abs(gl[a:b] - rg[a:b]).max() / max(1.0, np.abs(rg[a:b]).max())
        worst_block = max(worst_block, e_b)
        assert e_b <= 10 * TOL, name                       # every parameter block relative to ITS OWN largest gradient entry
    print("  worst per-block gradient error (relative to the block's own max) %.3g" % worst_block)
    net.loss_grad(torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(tp).cuda(), torch.from_numpy(st).cuda(),
                  keep_prob=keep_prob, seed=seed, sample_id0=sid0, iteration=it, loss_kind=loss_kind)
    assert np.array_equal(gl, net.gradloss.cpu().numpy().astype(np.float64))            # deterministic


def test_lstm_train_entry_point_learns():
    from reacherdistilation_b200 import lstm_train
    out = lstm_train.train(True, False, num_envs=48, iterations=100, verbose=False)
    assert len(out["losses"]) == 2 and np.isfinite(out["losses"]).all() and out["losses"][-1] < out["losses"][0]
    out["dataset"].close(); out["env"].close()


def test_graph_step_equals_loss_grad_then_adam():
    """rb_lstm_step (one CUDA-graph launch, device-side clock) vs rb_lstm_loss_grad + rb_adam_step with host counters."""
    from reacherdistilation_b200.student_nn import StudentLSTM
    a, b = StudentLSTM(seed=5), StudentLSTM(seed=5)
    B = 40
    ob, pp, tp, st = _data(B, 77)
    dob, dpp, dtp = torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(tp).cuda()
    for it in range(6):
        a.step(dob, dpp, dtp, None, keep_prob=0.5, seed=3)
        b.loss_grad(dob, dpp, dtp, None, keep_prob=0.5, seed=3, iteration=it)
        b.adam_step()
        if it == 0:
            assert torch.equal(a.gradloss, b.gradloss)                 # same kernels, same dropout mask (iteration from the device clock)
        assert torch.allclose(a.gradloss, b.gradloss, rtol=1e-3, atol=1e-3 * float(b.gradloss[:-1].abs().max())), it
        assert (a.params - b.params).abs().max().item() <= 1e-6
    assert a.t == b.t == 6


def test_lstm_loss_curve_matches_cpu_restatement():
    """`lstm_train.py:141-201` optimiser steps at the reference's batch (10 steps x 20 windows, keep_prob 0.5): 12 consecutive rb_lstm_step
    launches (one CUDA graph each, dropout iteration and Adam step from the device clock) on 12 different batches vs the float64 restatement
    with the same init, masks and TF-form Adam.  Stated tolerance: loss curve 1e-4 relative (measured 2.2e-6); parameters: a gradient element
    below the bf16x3 noise floor can take either sign, i.e. move by up to lr per step, so the gate is 99.9 % of the 511 880 parameters within
    1e-4 (measured 2.3e-6) and none further than steps * 2 * lr (measured 1.75e-3)."""
    from reacherdistilation_b200.student_nn import StudentLSTM
    B, steps, seed, sid0, lr = 20, 12, 11, 500, 1e-3
    net = StudentLSTM(seed=6, lr=lr, eps=1e-8)
    theta = net.params.cpu().numpy().astype(np.float64)
    opt = NN.AdamTF(theta.size, lr=lr, eps=1e-8)
    ids = np.arange(L.T * B, dtype=np.uint32) + sid0
    dev_losses, ref_losses = [], []
    for it in range(steps):
        ob, pp, tp, st = _data(B, 1000 + it)
        net.step(torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(tp).cuda(), torch.from_numpy(st).cuda(),
                 keep_prob=0.5, seed=seed, sample_id0=sid0)
        dev_losses.append(float(net.gradloss[-1]))
        obd = NN.student_input(ob.reshape(-1, 11), np.zeros((L.T * B, 4)), np.zeros(L.T * B), 0.5, seed, ids, it, dtype=np.float32)[:, :11]
        _, rl, rg = L.loss_grad(theta.astype(np.float32), obd.reshape(L.T, B, 11), pp, tp, st)
        theta = opt.update(theta, rg)
        ref_losses.append(rl)
    rel = np.abs(np.array(dev_losses) - np.array(ref_losses)) / np.maximum(1.0, np.abs(ref_losses))
    dp = np.abs(net.params.cpu().numpy() - theta)
    print("lstm loss curve: first %.5g last %.5g, max rel err %.3g; param err 99.9%% %.3g max %.3g" %
          (ref_losses[0], ref_losses[-1], rel.max(), np.quantile(dp, 0.999), dp.max()))
    assert ref_losses[-1] < ref_losses[0]
    assert rel.max() <= 1e-4
    assert np.quantile(dp, 0.999) <= 1e-4 and dp.max() <= steps * 2 * lr
    assert net.t == steps


def test_state_threaded_bptt_matches_oracle():
    """backup/lstm_bbpt.py:120-137: successive batches, the final (c, m) of one fed to the next as its initial state (value only), one Adam
    step per batch.  rb_lstm_loss_grad's final_state output + the threaded sequence vs the float64 restatement."""
    from reacherdistilation_b200.student_nn import StudentLSTM
    B, seed, sid0 = 20, 13, 40
    net = StudentLSTM(seed=8, lr=1e-3)
    theta = net.params.cpu().numpy().astype(np.float64)
    opt = NN.AdamTF(theta.size, lr=1e-3, eps=1e-8)
    ids = np.arange(L.T * B, dtype=np.uint32) + sid0
    s_dev, s_ref = net.zero_state(B), np.zeros((2, B, 200), np.float32)
    for it in range(4):
        ob, pp, tp, _ = _data(B, 300 + it)
        fin = torch.empty_like(s_dev)
        p_dev, s_in = net.params.cpu().numpy().copy(), s_dev.cpu().numpy().copy()
        net.loss_grad(torch.from_numpy(ob).cuda(), torch.from_numpy(pp).cuda(), torch.from_numpy(tp).cuda(), s_dev, keep_prob=0.5, seed=seed,
                      sample_id0=sid0, iteration=it, final_state_out=fin)
        dev_loss = float(net.gradloss[-1])
        net.adam_step()
        obd = NN.student_input(ob.reshape(-1, 11), np.zeros((L.T * B, 4)), np.zeros(L.T * B), 0.5, seed, ids, it, dtype=np.float32)[:, :11]
        obd = obd.reshape(L.T, B, 11)
        th32 = theta.astype(np.float32)
        _, rfin, _ = L.forward(th32, obd, pp, s_ref)
        _, rl, rg = L.loss_grad(th32, obd, pp, tp, s_ref)
        theta = opt.update(theta, rg)
        # kernel parity of the state output: the restatement run from the DEVICE's parameters and incoming state (1e-4, as the forward test);
        # chain parity: the independent float64 chain (own parameters, own threaded state) -- after the first Adam step the two parameter
        # vectors differ by up to 2 * lr on elements whose gradient is below the bf16x3 noise floor (test_lstm_loss_curve...), which the
        # state sees at the 1e-3 level and the loss at 1e-5
        e_k = np.abs(fin.cpu().numpy() - L.forward(p_dev, obd, pp, s_in)[1]).max()
        e_fin, e_l = np.abs(fin.cpu().numpy() - rfin).max(), abs(dev_loss - rl) / max(1.0, abs(rl))
        print("threaded batch %d: final state err %.3g (same parameters) / %.3g (independent chain, |state| max %.3g), loss %.5g rel err %.3g"
              % (it, e_k, e_fin, np.abs(rfin).max(), rl, e_l))
        assert e_k <= 2e-5 and e_fin <= 2.5e-3 and e_l <= 1e-4       # measured 4e-6 / 7.4e-4 (independent chain: see above) / 1e-5
        assert it == 0 or np.abs(rfin).max() > 1e-2                          # a non-trivial state is being threaded
        s_dev, s_ref = fin, np.asarray(rfin, np.float32)


def test_lstm_train_state_threaded_variant_runs():
    from reacherdistilation_b200 import lstm_train
    out = lstm_train.train(True, False, num_envs=48, iterations=100, verbose=False, thread_state=True, training_epochs=2)
    assert len(out["losses"]) == 2 and np.isfinite(out["losses"]).all() and out["losses"][-1] < out["losses"][0]
    assert out["student"].t == 200                                           # two optimiser steps per env step
    out["dataset"].close(); out["env"].close()
