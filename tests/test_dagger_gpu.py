"""DAgger loop (BASELINE config 4 semantics at test size): student loss curve vs the CPU restatement with identical teacher,
student init, seeds and dropout masks (loss-curve parity is defined against the restatement)."""
import numpy as np
import pytest
import torch

from oracle import nn_np as NN
from oracle import reacher_np as RN

pytestmark = pytest.mark.gpu


def _oracle_loop(kind, n, iters, seed, teacher_p, student_p, keep_prob, lr, eps, offset=0):
    from reacherdistilation_b200 import STUDENT_MLP
    env = RN.ReacherOracle(n, seed=seed, env_offset=offset)
    ob = env.reset()
    theta = student_p.astype(np.float64)
    opt = NN.AdamTF(theta.size, lr=lr, eps=eps)
    prev_t, prev_rec_rew, last_rew = np.zeros((n, 4)), np.zeros(n), np.zeros(n)
    losses, rews = [], []
    for it in range(iters):
        ob32 = ob.astype(np.float32)
        t = NN.policy_fwd(ob32, teacher_p).astype(np.float32)
        first = env.step_count == 0
        th32 = theta.astype(np.float32)
        if kind == STUDENT_MLP:
            pp = np.where(first[:, None], 0.0, prev_t)
            pr = np.where(first, 0.0, prev_rec_rew)
            x = NN.student_input(ob32, pp, pr, keep_prob, seed, env.env_ids, it)
            s, hs = NN.mlp_fwd(x, th32)
            l, ds = NN.kl_loss(s, t)
            g = NN.mlp_bwd(hs, th32, ds)
            if keep_prob < 1.0:            # dropout is for the training batch (mlp_train.py:151); the student ACTS on the clean observation (:171-186)
                s, _ = NN.mlp_fwd(NN.student_input(ob32, pp, pr, 1.0, seed, env.env_ids, it), th32)
        else:
            s = NN.policy_fwd(ob32, th32, nout=4)
            l, ds = NN.kl_loss(s, t)
            g = NN.policy_bwd(ob32, th32, ds)
        theta = opt.update(theta, g)
        ob, r, d = env.step(s[:, :2].astype(np.float32).astype(np.float64))
        prev_t, prev_rec_rew, last_rew = t.astype(np.float64), last_rew, r
        losses.append(l); rews.append(r.mean())
    return np.array(losses), np.array(rews), theta


@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
@pytest.mark.parametrize("kind_name,keep_prob", [("mlp", 0.5), ("mlp", 1.0), ("policy64", 1.0)])
def test_loss_curve_matches_cpu_restatement(kind_name, keep_prob, mode_name):
    from reacherdistilation_b200 import MODE_FP32, MODE_TC, STUDENT_MLP, STUDENT_POLICY64
    mode = MODE_TC if mode_name == "tc" else MODE_FP32
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    from reacherdistilation_b200.teacher import init_policy_params
    kind = STUDENT_MLP if kind_name == "mlp" else STUDENT_POLICY64
    n, iters, seed = 256, 70, 4
    tp = init_policy_params(seed=0, final_std=0.3)
    tr = DaggerTrainer(num_envs=n, seed=seed, student_kind=kind, keep_prob=keep_prob, teacher_params=tp, student_seed=1, lr=1e-3, eps=1e-8, mode=mode)
    sp = tr.student.params.cpu().numpy().copy()
    dev_losses, dev_rews = [], []
    for it in range(iters):
        tr.step()
        dev_losses.append(float(tr.last_loss())); dev_rews.append(float(tr.rew.mean()))
    ref_losses, ref_rews, theta = _oracle_loop(kind, n, iters, seed, tp, sp, keep_prob, 1e-3, 1e-8)
    rel = np.abs(np.array(dev_losses) - ref_losses) / np.maximum(1.0, np.abs(ref_losses))
    print("%s %s kp=%.1f loss curve: first %.4g last %.4g, max rel err %.3g; reward err %.3g" %
          (kind_name, mode_name, keep_prob, ref_losses[0], ref_losses[-1], rel.max(), np.abs(np.array(dev_rews) - ref_rews).max()))
    assert ref_losses[-1] < ref_losses[0]                 # it learns
    # stated loss-curve tolerance, closed loop over 70 iterations vs the float64 restatement, at ~4x the measured values: fp32 kernels 1.8e-5
    # -> 1e-4; tensor-core kernels (bf16x3 products, ~1e-5 per evaluation of teacher label AND student, then fed back) 2.8e-4 -> 1e-3
    tol = 1e-4 if mode_name == "fp32" else 1e-3
    assert rel.max() <= tol
    assert np.abs(np.array(dev_rews) - ref_rews).max() <= tol
    assert np.abs(tr.student.params.cpu().numpy() - theta).max() <= tol
    tr.close()


def test_train_entry_point_runs():
    from reacherdistilation_b200 import mlp_train
    out = mlp_train.train(True, False, num_envs=512, iterations=100, log_every=25, verbose=False)
    assert len(out["losses"]) == 4 and out["losses"][-1] < out["losses"][0]
    assert -1.0 < out["teacher_reward"] < 0.0
    out["trainer"].close()


@pytest.mark.parametrize("num_envs", [1, 48])
def test_reference_shaped_replay_loop(num_envs):
    """mlp_train.train_replay: the reference's loop shape (warm-up into the Dataset, window batches, test-batch acting)."""
    from reacherdistilation_b200 import mlp_train
    out = mlp_train.train_replay(True, False, num_envs=num_envs, iterations=150 if num_envs > 1 else 60, lr=1e-3, verbose=False,
                                 generations=64 if num_envs == 1 else 8)
    ds = out["dataset"]
    warm = -(-41 // num_envs) * num_envs                    # episodes recorded by the teacher before training starts (> 2 * MLP_BATCH_SIZE)
    assert ds.num_episodes() == warm + (150 // 50 if num_envs > 1 else 1) * num_envs
    assert len(out["losses"]) >= 1 and np.isfinite(out["losses"]).all()
    if num_envs > 1:
        assert out["losses"][-1] < out["losses"][0]
    ds.close(); out["env"].close()


@pytest.mark.parametrize("kind_name", ["mlp", "policy64"])
def test_graph_step_equals_three_launch_step(kind_name):
    """rb_dagger_step (one CUDA-graph launch, per-step values from the device-side clock) vs the three separate calls with host-side
    counters: same kernels, same arguments => identical losses, rewards and parameters (lr_t comes from a device pow(): 1e-6)."""
    from reacherdistilation_b200 import MODE_TC, STUDENT_MLP, STUDENT_POLICY64
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    kind = STUDENT_MLP if kind_name == "mlp" else STUDENT_POLICY64
    a = DaggerTrainer(num_envs=700, seed=2, student_kind=kind, mode=MODE_TC, lr=1e-3, use_graph=True)
    b = DaggerTrainer(num_envs=700, seed=2, student_kind=kind, mode=MODE_TC, lr=1e-3, use_graph=False)
    assert a.use_graph and not b.use_graph
    for it in range(60):
        a.step(); b.step()
        if it % 10 == 9:
            la, lb = float(a.last_loss()), float(b.last_loss())
            assert abs(la - lb) <= 1e-5 * max(1.0, abs(lb)), (it, la, lb)
            assert torch.equal(a.done, b.done) and (a.rew - b.rew).abs().max().item() <= 1e-5
    assert (a.student.params - b.student.params).abs().max().item() <= 1e-5
    assert a.iteration == b.iteration == 60 and a.student.t == b.student.t == 60
    a.use_graph = False                                    # mixed use: host counters stay authoritative
    a.step(); b.step()
    a.use_graph = True
    a.step(); b.step()
    assert abs(float(a.last_loss()) - float(b.last_loss())) <= 1e-5 * max(1.0, abs(float(b.last_loss())))
    a.close(); b.close()


def test_wait_loss_mailbox_matches_device_loss():
    """rb_dagger_wait_loss: the {loss, iteration} word the last kernel of the graph posts into mapped host memory is exactly the
    device-side loss of that iteration, every iteration, with no stream synchronise in between; an iteration that has already been
    overwritten is refused."""
    import ctypes as C
    from reacherdistilation_b200 import MODE_TC
    from reacherdistilation_b200._lib import ReacherB200Error, check, lib
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    tr = DaggerTrainer(num_envs=900, seed=5, mode=MODE_TC, lr=1e-3, use_graph=True)
    for it in range(25):
        tr.step()
        host = tr.wait_loss()                                  # polls; the stream is NOT synchronised here
        dev = float(tr.last_loss())                            # device read (synchronises)
        assert np.float32(host) == np.float32(dev), (it, host, dev)
    tr.step(); tr.step()
    torch.cuda.synchronize()
    out = C.c_float()
    with pytest.raises(ReacherB200Error):
        check(lib().rb_dagger_wait_loss(tr._h, tr.iteration - 1, C.byref(out)))
    assert np.float32(tr.wait_loss()) == np.float32(float(tr.last_loss()))
    tr.use_graph = False                                       # three-launch path: plain device read
    tr.step()
    assert np.isfinite(tr.wait_loss())
    tr.close()


@pytest.mark.parametrize("mode_name,use_graph", [("fp32", False), ("tc", True)])
def test_checkpoint_resume_is_bit_exact(mode_name, use_graph, tmp_path):
    """DaggerTrainer.state_dict / load_state_dict (rb_env_get/set_state, rb_dagger_get/set_state, student + Adam moments, iteration):
    a run restored into a FRESH trainer continues bit-identically to the uninterrupted one -- across an episode boundary (auto-reset from the
    restored Philox episode counter), with dropout (mask keyed by the restored iteration) and the `prev` / `prew` carry."""
    from reacherdistilation_b200 import MODE_FP32, MODE_TC, STUDENT_MLP
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    mode = MODE_TC if mode_name == "tc" else MODE_FP32
    kw = dict(num_envs=333, seed=6, student_kind=STUDENT_MLP, keep_prob=0.5, mode=mode, lr=1e-3, use_graph=use_graph, env_offset=1000)
    a = DaggerTrainer(**kw)
    for _ in range(37):
        a.step()
    path = str(tmp_path / "dagger.pt")
    torch.save(a.state_dict(), path)
    b = DaggerTrainer(**dict(kw, student_seed=99))              # different init: everything must come from the checkpoint
    b.load_state_dict(torch.load(path))
    assert b.iteration == 37 and b.student.t == 37
    for it in range(30):                                        # steps 37..66 of every env: crosses the reset at step 50
        a.step(); b.step()
        assert torch.equal(a.last_loss(), b.last_loss()), it
        assert torch.equal(a.rew, b.rew) and torch.equal(a.done, b.done) and torch.equal(a.x, b.x), it
    assert torch.equal(a.student.params, b.student.params) and torch.equal(a.student.m, b.student.m) and torch.equal(a.student.v, b.student.v)
    sa, sb = a.env.get_state(), b.env.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    with pytest.raises(ValueError):
        c = DaggerTrainer(**dict(kw, num_envs=64))
        try:
            c.load_state_dict(torch.load(path))
        finally:
            c.close()
    a.close(); b.close()


def test_train_entry_point_restore_continues_the_saved_run(tmp_path):
    """mlp_train.train(train, restore) (main.py:26-27 `-r`): 60 iterations, checkpoint, restore + 40 more == 100 iterations in one go."""
    from reacherdistilation_b200 import mlp_train
    ck = str(tmp_path / "student_mlp_b200.pt")
    full = mlp_train.train(True, False, num_envs=256, iterations=100, log_every=20, verbose=False)
    first = mlp_train.train(True, False, num_envs=256, iterations=60, log_every=20, verbose=False, checkpoint=ck)
    first["trainer"].close()
    rest = mlp_train.train(True, True, num_envs=256, iterations=40, log_every=20, verbose=False, checkpoint=ck)
    assert rest["teacher_reward"] is None                       # phase A is not replayed
    assert first["losses"] + rest["losses"] == full["losses"]
    assert torch.equal(rest["trainer"].student.params, full["trainer"].student.params)
    rest["trainer"].close(); full["trainer"].close()
