"""Philox4x32-10: Random123 known-answer vectors, numpy twin == C twin bit-for-bit, uniform mapping exactness."""
import numpy as np

from oracle import reacher_c as RC
from oracle import reacher_np as RN
from oracle.philox_np import philox4x32_10, philox4x32_10_raw, u32_to_uniform_f32

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10_raw(*ctr, *key)
        assert tuple(int(x) for x in got) == want
        seed = key[0] | (key[1] << 32)
        assert tuple(int(x) for x in RC.philox(seed, *ctr)) == want


def test_numpy_and_c_agree_on_random_counters():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2 ** 32, size=(4, 257), dtype=np.uint64).astype(np.uint32)
    seed = 0x1234567890ABCDEF
    got = philox4x32_10(seed, *c)
    for i in range(c.shape[1]):
        ref = RC.philox(seed, *(int(c[k, i]) for k in range(4)))
        assert all(int(got[k][i]) == int(ref[k]) for k in range(4))


def test_uniform_mapping_is_single_rounded_fma():
    x = np.array([0, 255, 256, 2 ** 31, 2 ** 32 - 1, 0x89ABCDEF], dtype=np.uint32)
    for lo, hi in ((-0.1, 0.1), (-0.2, 0.2), (-0.005, 0.005), (-1.0, 1.0)):
        u = u32_to_uniform_f32(x, lo, hi)
        assert u.dtype == np.float32
        assert (u >= np.float32(lo)).all() and (u < np.float32(hi)).all()
        assert u[0] == np.float32(lo) and u[1] == np.float32(lo)       # low 8 bits are dropped


def test_reset_streams_bit_exact_between_twins():
    n = 1000
    a = RN.ReacherOracle(n, seed=123, env_offset=77)
    b = RC.ReacherOracleC(n, seed=123, env_offset=77)
    a.reset(); b.reset()
    for k, name in enumerate(("q0", "q1", "v0", "v1", "tx", "ty")):
        assert np.array_equal(getattr(a, name), b.st[k]), name
        assert np.array_equal(getattr(a, name).astype(np.float32).astype(np.float64), getattr(a, name))   # float32-exact values
    act = RN.random_actions(123, a.env_ids, 5)
    assert act.dtype == np.float32 and np.abs(act).max() <= 1.0


def test_env_id_keying_is_partition_invariant():
    """Global env ids: a shard starting at offset k reproduces envs k.. of the unsharded run (multi-GPU invariance)."""
    full = RN.ReacherOracle(64, seed=9)
    part = RN.ReacherOracle(16, seed=9, env_offset=32)
    assert np.array_equal(full.reset()[32:48], part.reset())
