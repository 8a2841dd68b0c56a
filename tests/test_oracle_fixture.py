"""Pin the oracle against the reference's own recorded MuJoCo data.

tests/golden/reacher_fixture.npz = /root/reference src/distilation/tests/data/dataset.json re-encoded by
tests/golden/make_golden.py (25 episodes x 50 steps, float64)."""
import numpy as np
import pytest

from oracle import nn_np as NN
from oracle import reacher_c as RC
from oracle import reacher_np as RN


def _actions(fx):
    return np.where(fx["with_s"][..., None] == 1, fx["s"][..., :2], fx["t"][..., :2])


def _state_at(fx, k):
    ob = fx["ob"][:, k]
    q0, q1 = np.arctan2(ob[:, 2], ob[:, 0]), np.arctan2(ob[:, 3], ob[:, 1])
    return q0, q1, ob[:, 6], ob[:, 7], ob[:, 4], ob[:, 5], ob[:, 8] + ob[:, 4], ob[:, 9] + ob[:, 5]


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_kat1_2_3_one_step_transitions(fixture_data, impl):
    fx, act = fixture_data, _actions(fixture_data)
    E, T = fx["ob"].shape[:2]
    worst, worst_rew, limit_hits = 0.0, 0.0, 0
    for k in range(T - 1):
        st = _state_at(fx, k)
        if impl == "numpy":
            o = RN.ReacherOracle(E)
            o.set_state(*st)
        else:
            o = RC.ReacherOracleC(E)
            o.st[:] = np.stack(st)
            o.step_count[:] = 0
        nob, rew, _ = o.step(act[:, k], auto_reset=False)
        worst = max(worst, np.abs(nob - fx["ob"][:, k + 1]).max())
        worst_rew = max(worst_rew, np.abs(rew - fx["rew"][:, k + 1]).max())
        limit_hits += int((np.abs(st[1]) > 3.0).sum())
    assert worst < 1e-13, worst          # probe: 3.6e-15
    assert worst_rew < 5e-8, worst_rew   # json rounding of the stored reward: 1.1e-8
    assert limit_hits >= 30              # the joint-limit branch is exercised by the fixture (39 step-start states beyond +-3)


def test_kat4_open_loop_episode(fixture_data):
    fx, act = fixture_data, _actions(fixture_data)
    E, T = fx["ob"].shape[:2]
    o = RN.ReacherOracle(E)
    o.set_state(*_state_at(fx, 0))
    worst = 0.0
    for k in range(T - 1):
        nob, _, done = o.step(act[:, k], auto_reset=False)
        worst = max(worst, np.abs(nob - fx["ob"][:, k + 1]).max())
        assert not done.any()
    assert worst < 1e-12, worst          # probe: 3e-14
    _, _, done = o.step(act[:, T - 1], auto_reset=False)
    assert done.all()                    # TimeLimit: the 50th step ends the episode


def test_kat2_stale_kinematics_matters(fixture_data):
    """Using FK(final q) instead of the last-RK4-stage q must be visibly wrong (7e-4), i.e. the quirk is real."""
    fx = fixture_data
    ob = fx["ob"]
    q0, q1 = np.arctan2(ob[..., 2], ob[..., 0]), np.arctan2(ob[..., 3], ob[..., 1])
    px, py = RN.fk(q0, q1)
    err = np.abs(np.stack([px - ob[..., 4], py - ob[..., 5]], -1)[:, 1:] - ob[:, 1:, 8:10]).max()
    assert 1e-4 < err < 2e-3
    err0 = np.abs(np.stack([px - ob[..., 4], py - ob[..., 5]], -1)[:, 0] - ob[:, 0, 8:10]).max()
    assert err0 < 1e-15                  # KAT-5: at reset the kinematics are fresh


def test_kat5_reset_ranges_and_constants(fixture_data):
    ob = fixture_data["ob"]
    q0, q1 = np.arctan2(ob[:, 0, 2], ob[:, 0, 0]), np.arctan2(ob[:, 0, 3], ob[:, 0, 1])
    assert np.abs(q0).max() <= 0.1 and np.abs(q1).max() <= 0.1
    assert np.abs(ob[:, 0, 4:6]).max() <= 0.2 and np.abs(ob[:, 0, 6:8]).max() <= 0.005
    assert (ob[..., 10] == 0).all()
    assert (ob[:, :, 4:6] == ob[:, :1, 4:6]).all()      # target constant within an episode
    # the oracle's own Philox resets respect the same ranges
    o = RN.ReacherOracle(4096, seed=3)
    r = o.reset()
    assert np.abs(o.q0).max() <= 0.1 and np.abs(o.tx).max() <= 0.2 and np.abs(o.v0).max() <= 0.005 and (r[:, 10] == 0).all()


def test_kat6_prev_pdflat_shift(fixture_data):
    fx = fixture_data
    assert (fx["prev"][:, 0] == 0).all()
    teacher_eps = ~fx["with_s"].all(1)
    assert np.array_equal(fx["prev"][teacher_eps][:, 1:], fx["t"][teacher_eps][:, :-1])
    stud = fx["with_s"].all(1)
    assert np.array_equal(fx["prev"][stud][:, 1:], fx["s"][stud][:, :-1])     # old-format fixture


def test_kat8_kl_sums(fixture_data):
    want = {21: 6321.18735859, 22: 6451.91650283, 23: 5533.85279193, 24: 5462.02568571}
    for e, w in want.items():
        got, grad = NN.kl_loss(fixture_data["s"][e], fixture_data["t"][e])
        assert abs(got - w) < 1e-6
        assert grad.shape == (50, 4)


def test_teacher_logstd_constant(fixture_data):
    t = fixture_data["t"]
    assert np.allclose(t[..., 2], -3.2939295768737793) and np.allclose(t[..., 3], -3.3629262447357178)
