"""The fp32 arithmetic of the device physics, stated on the CPU: csrc/physics.cuh compiled for the host (tests/twin/, bit-exact twin of
the kernels -- tests/test_env_gpu.py checks that on the GPU box) against the float64 oracle at BASELINE config 2's full size
(4096 envs x 500 steps, Philox random actions, 10 auto-reset episodes per env).

Stated tolerance, per 50-step episode, |delta| / max(1, |x|) over the 11 observation entries and the reward:
  * episodes in free motion (|q1| never beyond 2.95):                        <= 2e-5   (measured 6.5e-6)
  * episodes that touch the soft joint limit (70 % of them with +-1 actions): <= 1.5e-4 (measured 1.1e-4; 99.9 % of them <= 6e-5)
    -- the contact is stiff (k = 2770, damping 105, impedance ramp over 1 mm): it multiplies the fp32 rounding of q1 and v1;
  * MuJoCo switches the limit constraint on discontinuously at dist = 0 (its damping term comes in at full strength), so an RK4 stage
    that lands within fp32 rounding distance of the threshold may take the other branch than the float64 trajectory: episodes in which
    the ORACLE sees a stage within 1e-6 rad of the threshold are set aside (about 30 of 40 960) and only bounded in number.
"""
import numpy as np

from oracle import reacher_c as RC
from oracle import reacher_np as RN
from twin.physics_twin import PhysicsTwin

TOL_FREE, TOL_LIMIT, GRAZE = 2e-5, 1.5e-4, 1e-6


def _run(n, T, seed):
    tw, orc = PhysicsTwin(n, seed=seed), RC.ReacherOracleC(n, seed=seed)
    o32, o64 = tw.reset(), orc.reset()
    assert np.array_equal(o32[:, 4:6].astype(np.float64), o64[:, 4:6])          # Philox reset draws are bit-exact (float32 values)
    ids = np.arange(n, dtype=np.uint32)
    err = np.zeros((T, n))
    q1max = np.zeros((T, n))
    graze = np.zeros((T, n))
    for t in range(T):
        act = RN.random_actions(seed, ids, t)
        ob, rw, dn = tw.step(act)
        q1max[t] = np.abs(orc.st[1])
        oref, rref, dref, gz = orc.step(act.astype(np.float64), graze=True)
        assert np.array_equal(dn, dref)
        e_ob = (np.abs(ob.astype(np.float64) - oref) / np.maximum(1.0, np.abs(oref))).max(axis=1)
        e_rw = np.abs(rw.astype(np.float64) - rref) / np.maximum(1.0, np.abs(rref))
        err[t], graze[t] = np.maximum(e_ob, e_rw), gz
        if dref.any():                                                      # auto-reset: the state is bit-exact again
            st = tw.get_state()
            assert np.array_equal(st["qpos"].astype(np.float64), orc.st[0:2].T) and np.array_equal(st["episode"], orc.episode)
            assert not st["qpos_lo"].any()
    return err, q1max, graze


def test_config2_full_size_fp32_twin_vs_float64_oracle():
    n, T = 4096, 500
    err, q1max, graze = _run(n, T, seed=0)
    E = err.reshape(T // 50, 50, n).max(axis=1).ravel()                     # per (episode, env)
    touch = q1max.reshape(T // 50, 50, n).max(axis=1).ravel() > 2.95
    grz = graze.reshape(T // 50, 50, n).min(axis=1).ravel() < GRAZE
    free, lim = E[~touch], E[touch & ~grz]
    print("config 2, fp32 twin vs float64 oracle: free-motion episodes %d: max %.2e | limit episodes %d: max %.2e, 99.9%% %.2e, median %.2e | "
          "set aside (stage within %.0e rad of the activation threshold): %d, worst of them %.2e"
          % (free.size, free.max(), lim.size, lim.max(), np.quantile(lim, 0.999), np.median(lim), GRAZE, int(grz.sum()), E[grz].max() if grz.any() else 0.0))
    assert free.max() <= TOL_FREE
    assert lim.max() <= TOL_LIMIT and np.quantile(lim, 0.999) <= 0.6 * TOL_LIMIT
    assert grz.sum() <= 2e-3 * E.size


def test_twin_rollout_equals_stepping_and_shards():
    """The fused-rollout loop and single steps are the same arithmetic; shards keyed by global env id reproduce the unsharded run."""
    n, T, seed = 96, 120, 3
    a, b = PhysicsTwin(n, seed=seed, env_offset=11), PhysicsTwin(n, seed=seed, env_offset=11)
    a.reset(); b.reset()
    out = a.rollout_random(T, step0=7)
    ids = (np.arange(n) + 11).astype(np.uint32)
    for t in range(T):
        act = RN.random_actions(seed, ids, 7 + t)
        assert np.array_equal(out["act"][t], act)
        ob, rw, dn = b.step(act)
        assert np.array_equal(out["obs"][t], ob) and np.array_equal(out["rew"][t], rw) and np.array_equal(out["done"][t].astype(bool), dn)
    part = PhysicsTwin(32, seed=seed, env_offset=11 + 64)
    part.reset()
    assert np.array_equal(part.rollout_random(T, step0=7)["obs"], out["obs"][:, 64:96])


def test_two_float_angles_round_trip_and_saturated_torque():
    """get_state / set_state carry the low parts (exact resume); a saturated torque for a whole episode (|v| -> 125 rad/s, stage increments
    of 1.3 rad: far outside what random actions or the teacher reach) stays within the limit-episode tolerance."""
    n = 64
    a, b = PhysicsTwin(n, seed=5), PhysicsTwin(n, seed=5)
    a.reset(); b.reset()
    a.rollout_random(23)
    st = a.get_state()
    assert np.abs(st["qpos_lo"]).max() > 0 and (np.abs(st["qpos_lo"]) <= np.spacing(np.abs(st["qpos"])) * 0.5 + 1e-30).all()
    b.set_state(**st)
    x, y = a.rollout_random(40, step0=23), b.rollout_random(40, step0=23)
    assert all(np.array_equal(x[k], y[k]) for k in x)
    tw, orc = PhysicsTwin(n, seed=9), RC.ReacherOracleC(n, seed=9)
    tw.reset(); orc.reset()
    act = np.tile(np.array([[1.5, -0.2]], np.float32), (n, 1))             # joint 0 saturated (ctrl clips at 1), joint 1 drifts
    worst = 0.0
    for t in range(49):
        ob, rw, dn = tw.step(act)
        oref, rref, dref = orc.step(act.astype(np.float64))
        worst = max(worst, float((np.abs(ob - oref) / np.maximum(1.0, np.abs(oref))).max()))
    assert np.abs(orc.st[2]).max() > 100.0                                   # really at ~125 rad/s
    print("saturated torque, 49 steps: |v0| reaches %.0f rad/s, worst error %.2e" % (np.abs(orc.st[2]).max(), worst))
    assert worst <= TOL_LIMIT
