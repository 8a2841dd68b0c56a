"""ORACLE (test infrastructure) -- ctypes wrapper over oracle/libreacher_oracle.so (reacher_oracle.c)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        L = C.CDLL(path)
        dp, ip, up, bp, fp = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                              C.POINTER(C.c_float))
        L.ro_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, up]
        L.ro_constants.argtypes = [dp]
        L.ro_max_threads.restype = C.c_int
        L.ro_reset_all.argtypes = [C.c_int, C.c_uint64, C.c_uint32, dp, ip, up, dp, C.c_int]
        L.ro_step.argtypes = [C.c_int, C.c_uint64, C.c_uint32, dp, ip, up, dp, dp, dp, bp, C.c_int, C.c_int]
        L.ro_step_graze.argtypes = [C.c_int, C.c_uint64, C.c_uint32, dp, ip, up, dp, dp, dp, bp, C.c_int, C.c_int, dp]
        L.ro_rollout_random.argtypes = [C.c_int, C.c_uint64, C.c_uint32, dp, ip, up, C.c_int, C.c_uint32, dp, C.c_int]
        L.ro_rollout_random.restype = C.c_double
        L.ro_policy_fwd.argtypes = [C.c_int, dp, fp, C.c_int, dp, C.c_int]
        L.ro_rollout_policy.argtypes = [C.c_int, C.c_uint64, C.c_uint32, dp, ip, up, C.c_int, fp, C.c_int, dp, dp, dp, bp, C.c_int]
        L.ro_rollout_policy.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def philox(seed, c0, c1, c2, c3):
    out = np.zeros(4, np.uint32)
    lib().ro_philox(seed, c0, c1, c2, c3, _p(out, C.c_uint32))
    return out


def constants():
    out = np.zeros(6)
    lib().ro_constants(_p(out, C.c_double))
    return out


def max_threads():
    return lib().ro_max_threads()


class ReacherOracleC:
    """Same surface as oracle.reacher_np.ReacherOracle, backed by the C restatement (OpenMP over envs)."""

    def __init__(self, num_envs, seed=0, env_offset=0, nthreads=0):
        self.n, self.seed, self.env_offset, self.nthreads = int(num_envs), int(seed), int(env_offset), int(nthreads)
        self.st = np.zeros((8, self.n))
        self.step_count = np.zeros(self.n, np.int32)
        self.episode = np.zeros(self.n, np.uint32)

    def reset(self):
        obs = np.zeros((self.n, 11))
        lib().ro_reset_all(self.n, self.seed, self.env_offset, _p(self.st, C.c_double), _p(self.step_count, C.c_int32),
                           _p(self.episode, C.c_uint32), _p(obs, C.c_double), self.nthreads)
        return obs

    def step(self, a, auto_reset=True, graze=False):
        """graze=True also returns, per env, the smallest | |q1| - 3 | over the 8 RK4 stage evaluations of this step: the distance of the
        trajectory from the joint-limit activation threshold, where the model itself is discontinuous."""
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(self.n, 2)
        obs, rew, done = np.zeros((self.n, 11)), np.zeros(self.n), np.zeros(self.n, np.uint8)
        gz = np.zeros(self.n) if graze else None
        lib().ro_step_graze(self.n, self.seed, self.env_offset, _p(self.st, C.c_double), _p(self.step_count, C.c_int32),
                            _p(self.episode, C.c_uint32), _p(a, C.c_double), _p(obs, C.c_double), _p(rew, C.c_double),
                            _p(done, C.c_uint8), int(auto_reset), self.nthreads, _p(gz, C.c_double))
        return (obs, rew, done.astype(bool), gz) if graze else (obs, rew, done.astype(bool))

    def rollout_random(self, T, step0=0, record=True):
        traj = np.zeros((T, self.n, 12)) if record else None
        tot = lib().ro_rollout_random(self.n, self.seed, self.env_offset, _p(self.st, C.c_double), _p(self.step_count, C.c_int32),
                                      _p(self.episode, C.c_uint32), T, step0, _p(traj, C.c_double), self.nthreads)
        return traj, tot

    def rollout_policy(self, T, params, nout=2, record=True):
        params = np.ascontiguousarray(params, dtype=np.float32)
        if record:
            ob, pd, rw, dn = np.zeros((T, self.n, 11)), np.zeros((T, self.n, 4)), np.zeros((T, self.n)), np.zeros((T, self.n), np.uint8)
        else:
            ob = pd = rw = dn = None
        tot = lib().ro_rollout_policy(self.n, self.seed, self.env_offset, _p(self.st, C.c_double), _p(self.step_count, C.c_int32),
                                      _p(self.episode, C.c_uint32), T, _p(params, C.c_float), nout, _p(ob, C.c_double),
                                      _p(pd, C.c_double), _p(rw, C.c_double), _p(dn, C.c_uint8), self.nthreads)
        return ob, pd, rw, dn, tot


def policy_fwd(obs, params, nout=2, nthreads=0):
    obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 11)
    params = np.ascontiguousarray(params, dtype=np.float32)
    out = np.zeros((obs.shape[0], 4))
    lib().ro_policy_fwd(obs.shape[0], _p(obs, C.c_double), _p(params, C.c_float), nout, _p(out, C.c_double), nthreads)
    return out
