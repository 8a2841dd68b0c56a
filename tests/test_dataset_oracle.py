"""Dataset semantics (reference dataset.py:118-290) on the CPU restatement, pinned by the reference's own fixture and by the
assertions of its unit test (src/distilation/tests/dataset_unit_test.py:13-94), re-expressed in pytest."""
import numpy as np

from oracle.dataset_np import STEPS_UNROLLED, DatasetOracle

T = STEPS_UNROLLED


def _load(fx, episodes):
    """Replay recorded fixture episodes through Dataset.write / flush (one env)."""
    ds = DatasetOracle(num_envs=1, generations=64)
    for e in episodes:
        for k in range(50):
            ds.write(fx["ob"][e, k][None], [fx["rew"][e, k]], fx["t"][e, k][None], fx["s"][e, k][None], "s" if fx["with_s"][e, k] else "t")
        ds.flush()
    return ds


def test_prev_is_teacher_pdflat_of_previous_record(fixture_data):
    """dataset_unit_test.py:13-18 and SURVEY KAT 6: prev[0] == 0, prev[k] == t[k-1] -- checked against the RECORDED prev on the
    teacher-stepped episodes (the old-format fixture stores the student pdflat on student episodes)."""
    fx = fixture_data
    ds = _load(fx, range(21))
    for e in range(21):
        ep = ds.data_in_memory[e]
        assert np.array_equal(ep[0]["prev"], np.zeros(4)) and ep[0]["prew"] == 0.0
        for k in range(1, 50):
            assert np.array_equal(ep[k]["prev"], ds.pdflat_at(ep, k - 1))
            assert np.array_equal(ep[k]["prev"], fx["prev"][e, k])            # the reference's own recording
            assert ep[k]["prew"] == ep[k - 1]["rew"]
    assert ds.num_total_episodes == 21


def test_ob_batch_test_array_three_regimes(fixture_data):
    """dataset_unit_test.py:46-94: episode shorter than / equal to / longer than STEPS_UNROLLED-1."""
    fx = fixture_data
    ob = np.full((1, 11), -10.0)
    for length in (T - 3, T - 1, T + 5, 0, 49):
        ds = DatasetOracle(num_envs=1)
        for k in range(length):
            ds.write(fx["ob"][0, k][None], [fx["rew"][0, k]], fx["t"][0, k][None])
        o, p, w = ds.test_batch(ob)
        assert o.shape == (T, 1, 11) and p.shape == (T, 1, 4) and w.shape == (T, 1, 1)
        pad = max(0, T - 1 - length)
        assert np.array_equal(o[:pad, 0], np.zeros((pad, 11)))
        first = max(0, length - (T - 1))
        assert np.array_equal(o[pad:T - 1, 0], fx["ob"][0, first:length])
        assert np.array_equal(o[T - 1, 0], ob[0])
        # dataset_unit_test.py:21-26: rows i < T-1 are the `prev` fields of the last T-1 records, the last row is t of the last record
        for i in range(T - 1):
            j = length - T + 1 + i
            assert np.array_equal(p[i, 0], ds.curr[0][j]["prev"] if j >= 0 else np.zeros(4))
            assert w[i, 0, 0] == (ds.curr[0][j]["prew"] if j >= 0 else 0.0)
        assert np.array_equal(p[T - 1, 0], ds.pdflat_at(ds.curr[0], length - 1))
        assert w[T - 1, 0, 0] == ds.rew_at(ds.curr[0], length - 1)


def test_training_batch_windows(fixture_data):
    """dataset.py:184-202: B episodes with replacement, one shared start in [0, 40], time-major [T,B,.] windows."""
    fx = fixture_data
    ds = _load(fx, range(25))
    seen_starts = set()
    for draw in range(40):
        ob, t, prev, prew, eps, start = ds.training_batch(draw)
        assert ob.shape == (T, 20, 11) and t.shape == (T, 20, 4) and prev.shape == (T, 20, 4) and prew.shape == (T, 20, 1)
        assert 0 <= start <= 50 - T and (eps >= 0).all() and (eps < 25).all()
        seen_starts.add(start)
        for b, e in enumerate(eps):
            assert np.array_equal(ob[:, b], fx["ob"][e, start:start + T])
            assert np.array_equal(t[:, b], fx["t"][e, start:start + T])
            if start > 0:
                assert np.array_equal(prev[:, b], fx["t"][e, start - 1:start + T - 1])
    assert len(seen_starts) > 10
    a, b = ds.training_batch(7), ds.training_batch(7)
    assert np.array_equal(a[0], b[0]) and a[5] == b[5]                       # deterministic given (seed, draw)
