"""Device env kernels vs the float64 oracle (and vs the reference's recorded MuJoCo data) through the C ABI.

Stated tolerance (fp32 state; tests/test_physics_twin_cpu.py derives it on the CPU from the host build of the same source): per
50-step episode |delta| <= 1.5e-4 * max(1, |x|) on obs / reward (2e-5 for episodes that stay off the soft joint limit; episodes in
which the oracle sees an RK4 stage within 1e-6 rad of the limit's discontinuous activation threshold are set aside); resets and
RNG-driven indexing bit-exact.  The kernels are also BIT-IDENTICAL to that host build (tests/twin/): state and observations to the
bit, the reward to the last ulp of the MUFU square root."""
import numpy as np
import pytest
import torch

from oracle import reacher_c as RC
from oracle import reacher_np as RN

pytestmark = pytest.mark.gpu
TOL, TOL_FREE, GRAZE = 1.5e-4, 2e-5, 1e-6


def _close(dev, ref, tol=TOL):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(np.asarray(dev, dtype=np.float64) - ref) / np.maximum(1.0, np.abs(ref))
    return float(err.max())


def _mk(n, **kw):
    from reacherdistilation_b200.env import VecReacher
    return VecReacher(num_envs=n, **kw)


@pytest.mark.parametrize("n,offset", [(1, 0), (31, 5), (4096, 0), (1000, 4096)])
def test_reset_bit_exact(n, offset):
    env = _mk(n, seed=42, env_offset=offset)
    obs = env.reset().cpu().numpy()
    st = env.get_state()
    o = RN.ReacherOracle(n, seed=42, env_offset=offset)
    ref = o.reset()
    assert np.array_equal(st["qpos"].cpu().numpy().astype(np.float64), np.stack([o.q0, o.q1], -1))
    assert np.array_equal(st["qvel"].cpu().numpy().astype(np.float64), np.stack([o.v0, o.v1], -1))
    assert np.array_equal(st["target"].cpu().numpy().astype(np.float64), np.stack([o.tx, o.ty], -1))
    assert (st["step"].cpu().numpy() == 0).all() and (st["episode"].cpu().numpy() == 0).all()
    assert _close(obs, ref, 1e-6) < 1e-6
    assert (obs[:, 10] == 0).all()
    env.close()


def test_config2_4096_envs_500_steps_random_actions():
    """BASELINE.json config 2: lock-step parity over 500 steps (10 auto-reset episodes) with Philox random actions -- against the float64
    oracle within the stated tolerance, and against the host build of csrc/physics.cuh bit for bit."""
    from twin.physics_twin import PhysicsTwin
    n, T, seed = 4096, 500, 0
    env = _mk(n, seed=seed)
    o = RC.ReacherOracleC(n, seed=seed)
    tw = PhysicsTwin(n, seed=seed)
    obs = env.reset()
    o.reset()
    assert np.array_equal(obs.cpu().numpy(), tw.reset())
    err, q1max, graze = np.zeros((T, n)), np.zeros((T, n)), np.zeros((T, n))
    ids = np.arange(n, dtype=np.uint32)
    for t in range(T):
        act = RN.random_actions(seed, ids, t)                      # float32, identical bits on all three sides
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        ob_h, rew_h = obs.cpu().numpy(), rew.cpu().numpy()
        q1max[t] = np.abs(o.st[1])
        oref, rref, dref, graze[t] = o.step(act.astype(np.float64), graze=True)
        ob_t, rew_t, done_t = tw.step(act)
        assert np.array_equal(ob_h, ob_t), "device observation differs from the host build of the same arithmetic at step %d" % t
        assert np.abs(rew_h - rew_t).max() <= 5e-7 and np.array_equal(done.cpu().numpy().astype(bool), done_t)
        assert np.array_equal(done_t, dref)
        err[t] = np.maximum((np.abs(ob_h.astype(np.float64) - oref) / np.maximum(1.0, np.abs(oref))).max(axis=1),
                            np.abs(rew_h.astype(np.float64) - rref) / np.maximum(1.0, np.abs(rref)))
        if dref.any():                                             # auto-reset: state bit-exact again => drift does not carry over
            st = env.get_state()
            assert np.array_equal(st["qpos"].cpu().numpy().astype(np.float64), o.st[0:2].T)
            assert np.array_equal(st["qvel"].cpu().numpy().astype(np.float64), o.st[2:4].T)
            assert np.array_equal(st["target"].cpu().numpy().astype(np.float64), o.st[4:6].T)
            assert np.array_equal(st["episode"].cpu().numpy().astype(np.uint32), o.episode)
    st, stt = env.get_state(), tw.get_state()
    for k in ("qpos", "qpos_lo", "qvel", "target", "fingertip"):
        assert np.array_equal(st[k].cpu().numpy(), stt[k]), k
    E = err.reshape(T // 50, 50, n).max(axis=1).ravel()
    touch = q1max.reshape(T // 50, 50, n).max(axis=1).ravel() > 2.95
    grz = graze.reshape(T // 50, 50, n).min(axis=1).ravel() < GRAZE
    free, lim = E[~touch], E[touch & ~grz]
    print("config2 parity: bit-identical to the host twin; vs float64 oracle: free-motion episodes max %.3g (tol %.2g), joint-limit episodes max %.3g "
          "(tol %.2g), %d of %d episodes set aside (RK4 stage within %.0e rad of the limit's activation threshold)"
          % (free.max(), TOL_FREE, lim.max(), TOL, int(grz.sum()), E.size, GRAZE))
    assert free.max() <= TOL_FREE and lim.max() <= TOL and grz.sum() <= 2e-3 * E.size
    env.close()


def test_fused_random_rollout_equals_stepping():
    """The fused T-step kernel and T single-step launches are the same arithmetic: bit-identical trajectories."""
    n, T, seed = 777, 120, 3
    a, b = _mk(n, seed=seed, env_offset=11), _mk(n, seed=seed, env_offset=11)
    a.reset(); b.reset()
    out = a.rollout_random(T, step0=7, record_obs=True, record_act=True)
    ids = (np.arange(n) + 11).astype(np.uint32)
    for t in range(T):
        act = torch.from_numpy(RN.random_actions(seed, ids, 7 + t)).cuda()
        assert torch.equal(out["act"][t], act)
        obs, rew, done, _ = b.step(act)
        assert torch.equal(out["obs"][t], obs) and torch.equal(out["rew"][t], rew) and torch.equal(out["done"][t], done)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    a.close(); b.close()


def test_device_vs_reference_mujoco_fixture(fixture_data):
    """One-step transitions of the reference's own recorded MuJoCo episodes, on the device, from identical states/actions."""
    fx = fixture_data
    ob = fx["ob"]
    act = np.where(fx["with_s"][..., None] == 1, fx["s"][..., :2], fx["t"][..., :2])
    E, T = ob.shape[:2]
    cur, nxt, a = ob[:, :-1].reshape(-1, 11), ob[:, 1:].reshape(-1, 11), act[:, :-1].reshape(-1, 2)
    n = cur.shape[0]
    env = _mk(n, seed=0)
    env.reset()
    q = np.stack([np.arctan2(cur[:, 2], cur[:, 0]), np.arctan2(cur[:, 3], cur[:, 1])], -1)
    env.set_state(qpos=q, qvel=cur[:, 6:8], target=cur[:, 4:6], fingertip=cur[:, 8:10] + cur[:, 4:6], step=np.zeros(n, np.int32))
    obs, rew, done, _ = env.step(torch.from_numpy(a.astype(np.float32)).cuda())
    e_obs = _close(obs.cpu().numpy(), nxt)
    e_rew = _close(rew.cpu().numpy(), fx["rew"][:, 1:].reshape(-1))
    print("device vs recorded MuJoCo: one-step obs err %.3g, reward err %.3g" % (e_obs, e_rew))
    assert e_obs <= 2e-5 and e_rew <= 2e-6      # single step: fp32 rounding only
    assert not done.any()
    env.close()


def test_open_loop_fixture_episodes_on_device(fixture_data):
    """49-step open-loop replay of the 25 recorded episodes (includes 6 joint-limit episodes) within the stated tolerance."""
    fx = fixture_data
    ob = fx["ob"]
    act = np.where(fx["with_s"][..., None] == 1, fx["s"][..., :2], fx["t"][..., :2]).astype(np.float32)
    E, T = ob.shape[:2]
    env = _mk(E, seed=0)
    env.reset()
    o0 = ob[:, 0]
    env.set_state(qpos=np.stack([np.arctan2(o0[:, 2], o0[:, 0]), np.arctan2(o0[:, 3], o0[:, 1])], -1), qvel=o0[:, 6:8], target=o0[:, 4:6],
                  fingertip=o0[:, 8:10] + o0[:, 4:6], step=np.zeros(E, np.int32))
    worst = 0.0
    for k in range(T - 1):
        obs, rew, done, _ = env.step(torch.from_numpy(act[:, k]).cuda())
        worst = max(worst, _close(obs.cpu().numpy(), ob[:, k + 1]))
    print("open-loop 49 steps vs recorded MuJoCo: worst err %.3g" % worst)
    assert worst <= 4e-5                        # measured 1.1e-5 (6 of the 25 episodes touch the joint limit)
    env.close()


def test_host_surface_single_env_gym_protocol():
    from reacherdistilation_b200.env import make_mujoco_env
    env = make_mujoco_env("Reacher-v2", 0)
    assert env.observation_space.shape == (11,) and env.action_space.shape == (2,)
    ob = env.reset()
    assert isinstance(ob, np.ndarray) and ob.shape == (11,)
    o = RN.ReacherOracle(1, seed=0)
    ref = o.reset()[0]
    assert np.abs(ob - ref).max() < 1e-6
    dones = 0
    for t in range(100):
        a = np.array([[0.3, -0.2]], np.float32)
        ob, r, new, info = env.step(a)
        oref, rref, dref = o.step(a.astype(np.float64))
        assert ob.shape == (11,) and isinstance(r, float) and isinstance(new, bool)
        assert _close(ob, oref[0]) <= TOL and abs(r - rref[0]) <= TOL and new == bool(dref[0])
        dones += new
    assert dones == 2
    env.close()


def test_multi_gpu_invariance_by_global_env_id():
    """Shards keyed by global env id reproduce the unsharded trajectories bit-for-bit (what 1/2/4/8-GPU runs rely on)."""
    n, T = 512, 60
    full = _mk(n, seed=5)
    full.reset()
    ref = full.rollout_random(T)
    for lo, hi in ((0, 128), (128, 512)):
        part = _mk(hi - lo, seed=5, env_offset=lo)
        part.reset()
        out = part.rollout_random(T)
        assert torch.equal(out["obs"], ref["obs"][:, lo:hi]) and torch.equal(out["rew"], ref["rew"][:, lo:hi])
        part.close()
    full.close()


def test_ragged_sizes_and_empty_rollout():
    for n in (1, 2, 33, 127, 129):
        env = _mk(n, seed=1)
        env.reset()
        out = env.rollout_random(3)
        assert out["obs"].shape == (3, n, 11) and torch.isfinite(out["obs"]).all()
        assert env.rollout_random(0)["rew"].numel() == 0
        env.close()
