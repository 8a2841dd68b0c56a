"""ORACLE (test infrastructure) -- Philox4x32-10 counter-based RNG in numpy (Salmon et al., SC'11; Random123).

north_star asks for Philox-based reset / target sampling that is bit-exact given the seed; the reference's own
stream (gym seeding -> numpy MT19937, call site /root/reference src/distilation/mlp_train.py:21 `make_mujoco_env(.., 0)`)
cannot be reproduced (gym absent) -- "parity unpinned" for the reset STREAM by design; the generator itself is pinned
by the Random123 known-answer vectors in tests/test_philox.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10_raw(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox4x32_10(seed, c0, c1, c2, c3):
    """key = (seed & 0xffffffff, seed >> 32)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10_raw(c0, c1, c2, c3, seed & 0xFFFFFFFF, seed >> 32)


def u32_to_uniform_f32(x, lo, hi):
    """u = (x >> 8) * 2^-24 in [0,1); result = fmaf(u, hi - lo, lo) in float32 (single rounding).

    float64 evaluation of u*w + lo is exact here (<= 48 significant bits), so one cast reproduces fmaf bit-for-bit."""
    lo, hi = np.float32(lo), np.float32(hi)
    w = np.float32(hi - lo)
    u = (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * (2.0 ** -24)
    return (u * np.float64(w) + np.float64(lo)).astype(np.float32)
