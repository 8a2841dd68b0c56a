"""numpy and C restatements agree (different libm / evaluation order => 1e-12), including auto-reset and policy rollout."""
import numpy as np

from oracle import nn_np as NN
from oracle import reacher_c as RC
from oracle import reacher_np as RN


def _teacher(seed=0, nout=2):
    rng = np.random.default_rng(seed)
    return np.concatenate([np.zeros(11), np.ones(11), NN.normc_init(rng, (11, 64), 1.0).ravel(), np.zeros(64),
                           NN.normc_init(rng, (64, 64), 1.0).ravel(), np.zeros(64), NN.normc_init(rng, (64, nout), 0.01 if nout == 2 else 0.3).ravel(),
                           np.zeros(nout), [-3.29, -3.36]]).astype(np.float32)


def test_step_twins_with_auto_reset():
    n = 512
    a, b = RN.ReacherOracle(n, seed=7, env_offset=100), RC.ReacherOracleC(n, seed=7, env_offset=100)
    assert np.array_equal(a.reset(), b.reset())
    worst = 0.0
    for t in range(120):
        act = RN.random_actions(7, a.env_ids, t).astype(np.float64)
        oa, ra, da = a.step(act)
        ob, rb, db = b.step(act)
        assert np.array_equal(da, db)
        worst = max(worst, np.abs(oa - ob).max(), np.abs(ra - rb).max())
    assert worst < 1e-11
    assert (a.episode == 2).all() and np.array_equal(a.episode, b.episode)


def test_rollout_random_matches_stepping():
    n = 64
    b1, b2 = RC.ReacherOracleC(n, seed=5), RC.ReacherOracleC(n, seed=5)
    b1.reset(); b2.reset()
    traj, _ = b1.rollout_random(70, step0=3)
    ids = np.arange(n, dtype=np.uint32)
    for t in range(70):
        ob, rw, _ = b2.step(RN.random_actions(5, ids, 3 + t).astype(np.float64))
        assert np.array_equal(traj[t, :, :11], ob) and np.array_equal(traj[t, :, 11], rw)


def test_policy_twins_and_rollout():
    p = _teacher()
    rng = np.random.default_rng(1)
    ob = rng.standard_normal((300, 11)) * 3
    assert np.abs(NN.policy_fwd(ob, p) - RC.policy_fwd(ob, p)).max() < 1e-14
    p4 = _teacher(2, nout=4)
    assert np.abs(NN.policy_fwd(ob, p4, nout=4) - RC.policy_fwd(ob, p4, nout=4)).max() < 1e-14
    n, T = 32, 60
    c = RC.ReacherOracleC(n, seed=11); c.reset()
    obs, pd, rew, done, _ = c.rollout_policy(T, p)
    a = RN.ReacherOracle(n, seed=11); o = a.reset()
    for t in range(T):
        assert np.abs(o - obs[t]).max() < 1e-11
        flat = NN.policy_fwd(o, p)
        assert np.abs(flat - pd[t]).max() < 1e-11
        o, r, d = a.step(flat[:, :2].astype(np.float32).astype(np.float64))
        assert np.abs(r - rew[t]).max() < 1e-11 and np.array_equal(d, done[t].astype(bool))
    assert done[49].all() and not done[48].any()
