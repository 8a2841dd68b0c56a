"""BASELINE.json configs at their FULL sizes, through size-independent properties (the oracle needs minutes at these sizes):

config 3 -- 65 536 envs, fused teacher, 50-step chunks (mlp_train.py:120-139 batched): every buffer row is self-consistent
            (reward = -|fingertip - target| - |a|^2 from the SAME row's obs and pdflat, reacher.py semantics pinned in SURVEY App. A;
            cos^2 + sin^2 = 1; 11th obs = 0; target constant inside an episode), `done` is exactly the TimeLimit(50) pattern, the
            post-reset rows are bit-identical to the Philox reset oracle, the recorded pdflat is the teacher of the recorded obs
            (checked with the oracle on a seeded subsample), two shards keyed by global env id reproduce the unsharded buffers
            bit for bit, and a second run from the same seed is bit-identical;
config 4 -- one DAgger iteration on 262 144 envs (one GPU holds the whole batch): loss and gradient equal the sum over 8 shards
            of 32 768 envs computed separately (what the 8-GPU all-reduce produces: KL is a sum, loss.py:11), the fixed-order
            reduction makes a repeat bit-identical, and the loss on a seeded subsample of rows equals the float64 restatement.
"""
import numpy as np
import pytest
import torch

from oracle import nn_np as NN
from oracle import reacher_np as RN

pytestmark = pytest.mark.gpu


def _tc():
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200._lib import lib
    return (MODE_TC, 5e-5) if lib().rb_mode_available(MODE_TC) else (MODE_FP32, 2e-6)


def test_config3_65536_envs_fused_teacher_chunk_properties():
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    mode, ptol = _tc()
    n, T, seed = 65536, 50, 0
    p = init_policy_params(seed=0, final_std=0.3)
    p_dev = torch.from_numpy(p).cuda()
    env = VecReacher(num_envs=n, seed=seed)
    ob0 = env.reset().clone()
    out = {k: v.clone() for k, v in env.rollout_policy(p_dev, T, nout=2, mode=mode).items()}
    out2 = env.rollout_policy(p_dev, 2, nout=2, mode=mode)           # first rows of the next chunk = the auto-reset observations
    obs, pd, rew, done = out["obs"], out["pdflat"], out["rew"], out["done"]
    assert obs.shape == (T, n, 11) and pd.shape == (T, n, 4) and rew.shape == (T, n) and done.shape == (T, n)
    assert torch.equal(obs[0], ob0)
    # TimeLimit(50): done exactly on the last row of the chunk, for every env
    assert int(done[:-1].sum()) == 0 and int(done[-1].sum()) == n
    # row self-consistency (fp32 roundings only: sqrt.approx + 3 FMAs)
    dist = torch.sqrt(obs[..., 8] ** 2 + obs[..., 9] ** 2)
    r_ref = -dist - (pd[..., 0] ** 2 + pd[..., 1] ** 2)
    assert float((rew - r_ref).abs().max()) <= 2e-6
    assert float((obs[..., 0] ** 2 + obs[..., 2] ** 2 - 1).abs().max()) <= 1e-6
    assert float((obs[..., 1] ** 2 + obs[..., 3] ** 2 - 1).abs().max()) <= 1e-6
    assert float(obs[..., 10].abs().max()) == 0.0
    assert torch.equal(obs[1:, :, 4:6], obs[:-1, :, 4:6])           # target fixed inside the episode
    assert torch.equal(pd[..., 2:], p_dev[-2:].expand(T, n, 2))     # logstd half of pdflat = the policy's logstd variable
    assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
    # resets: episode 1 of every env, bit-exact against the Philox oracle (RNG-driven resets are bit-exact given the seed)
    ids = np.arange(n, dtype=np.uint32)
    q0, q1, v0, v1, tx, ty = RN.reset_draws(seed, ids, np.ones(n, np.uint32))
    ob1 = out2["obs"][0].cpu().numpy()
    assert np.array_equal(ob1[:, 4], tx.astype(np.float32)) and np.array_equal(ob1[:, 5], ty.astype(np.float32))
    assert np.array_equal(ob1[:, 6], v0.astype(np.float32)) and np.array_equal(ob1[:, 7], v1.astype(np.float32))
    assert np.abs(ob1[:, 0] - np.cos(q0)).max() <= 2e-7 and np.abs(ob1[:, 3] - np.sin(q1)).max() <= 2e-7
    # teacher-forced on a seeded subsample of rows
    idx = np.random.default_rng(0).integers(0, T * n, 20000)
    ob_s = obs.reshape(-1, 11)[torch.from_numpy(idx).cuda()].cpu().numpy()
    pd_s = pd.reshape(-1, 4)[torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.abs(pd_s - NN.policy_fwd(ob_s, p)).max() <= ptol
    # sharding by global env id (what the 2/4/8-GPU runs do) and run-to-run determinism, bit for bit
    for lo, hi in ((0, 8192), (8192, 65536)):
        part = VecReacher(num_envs=hi - lo, seed=seed, env_offset=lo)
        part.reset()
        po = part.rollout_policy(p_dev, T, nout=2, mode=mode)
        for k in ("obs", "pdflat", "rew", "done"):
            assert torch.equal(po[k], out[k][:, lo:hi]), (k, lo, hi)
        part.close()
    env.close()


def test_config4_262144_env_dagger_iteration_equals_sum_of_8_shards():
    from reacherdistilation_b200 import STUDENT_MLP
    from reacherdistilation_b200.mlp_train import DaggerTrainer
    from reacherdistilation_b200.teacher import init_policy_params
    mode, ptol = _tc()
    n, shards, seed = 262144, 8, 3
    tp = init_policy_params(seed=0, final_std=0.3)
    full = DaggerTrainer(num_envs=n, seed=seed, student_kind=STUDENT_MLP, keep_prob=0.5, teacher_params=tp, student_seed=1, lr=1e-4, mode=mode)
    P0 = full.student.params.clone()
    full.step()
    torch.cuda.synchronize()
    gl_full = full.student.gradloss.clone().double()
    x_full, t_full = full.x.clone(), full.t_pd.clone()
    # the float64 restatement on a seeded subsample of the rows of this very batch
    idx = np.random.default_rng(1).integers(0, n, 4096)
    xs, ts = x_full[torch.from_numpy(idx).cuda()].cpu().numpy(), t_full[torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.abs(ts - NN.policy_fwd(full.obs[torch.from_numpy(idx).cuda()].cpu().numpy(), tp)).max() <= ptol
    acc = torch.zeros_like(gl_full)
    per = n // shards
    for r in range(shards):
        tr = DaggerTrainer(num_envs=per, seed=seed, student_kind=STUDENT_MLP, keep_prob=0.5, teacher_params=tp, student_seed=1, lr=1e-4, mode=mode,
                           env_offset=r * per)
        assert torch.equal(tr.student.params, P0)
        tr.step()
        torch.cuda.synchronize()
        assert torch.equal(tr.x, x_full[r * per:(r + 1) * per])          # same envs, same dropout masks: keyed by GLOBAL env id
        assert torch.equal(tr.t_pd, t_full[r * per:(r + 1) * per])
        acc += tr.student.gradloss.double()
        tr.close()
    scale = float(gl_full[:-1].abs().max())
    # different summation trees (one grid over 2048 tiles vs 8 x 256 tiles): fp32 reduction-order noise only
    assert float((acc[:-1] - gl_full[:-1]).abs().max()) <= 2e-5 * scale
    assert abs(float(acc[-1] - gl_full[-1])) <= 2e-6 * abs(float(gl_full[-1]))
    s, hs = NN.mlp_fwd(xs, P0.cpu().numpy())
    l, _ = NN.kl_loss(s, ts)
    full.close()
    # loss of the subsample computed by the device on exactly those rows
    from reacherdistilation_b200.student_nn import StudentNet
    st = StudentNet(STUDENT_MLP, seed=1, mode=mode, params=P0.cpu().numpy())
    st.loss_grad(torch.from_numpy(xs).cuda(), torch.from_numpy(ts).cuda())
    torch.cuda.synchronize()
    assert abs(float(st.gradloss[-1]) - l) <= 2e-3 * abs(l)
