"""TEST INFRASTRUCTURE: ctypes wrapper of the host build of csrc/physics.cuh (see physics_twin.cu)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        vp = C.c_void_p
        L.twin_create.restype = vp
        L.twin_create.argtypes = [C.c_int64, C.c_uint64, C.c_uint32]
        L.twin_destroy.argtypes = [vp]
        L.twin_reset.argtypes = [vp, vp]
        L.twin_step.argtypes = [vp, vp, vp, vp, vp]
        L.twin_rollout_random.argtypes = [vp, C.c_int, C.c_uint32, vp, vp, vp, vp]
        L.twin_get_state.argtypes = [vp] + [vp] * 7
        L.twin_set_state.argtypes = [vp] + [vp] * 7
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class PhysicsTwin:
    """fp32 host twin of the device env kernels: same surface as oracle.reacher_np.ReacherOracle, float32 arrays."""

    def __init__(self, num_envs, seed=0, env_offset=0):
        self.n = int(num_envs)
        self._h = lib().twin_create(self.n, int(seed), int(env_offset))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().twin_destroy(self._h)
            self._h = None

    def reset(self):
        obs = np.empty((self.n, 11), np.float32)
        lib().twin_reset(self._h, _p(obs))
        return obs

    def step(self, act):
        act = np.ascontiguousarray(act, np.float32).reshape(self.n, 2)
        obs, rew, done = np.empty((self.n, 11), np.float32), np.empty(self.n, np.float32), np.empty(self.n, np.uint8)
        lib().twin_step(self._h, _p(act), _p(obs), _p(rew), _p(done))
        return obs, rew, done.astype(bool)

    def rollout_random(self, T, step0=0):
        n = self.n
        out = dict(obs=np.empty((T, n, 11), np.float32), act=np.empty((T, n, 2), np.float32), rew=np.empty((T, n), np.float32),
                   done=np.empty((T, n), np.uint8))
        lib().twin_rollout_random(self._h, T, step0, _p(out["obs"]), _p(out["act"]), _p(out["rew"]), _p(out["done"]))
        return out

    def get_state(self):
        n = self.n
        out = dict(qpos=np.empty((n, 2), np.float32), qvel=np.empty((n, 2), np.float32), target=np.empty((n, 2), np.float32),
                   fingertip=np.empty((n, 2), np.float32), step=np.empty(n, np.int32), episode=np.empty(n, np.uint32),
                   qpos_lo=np.empty((n, 2), np.float32))
        lib().twin_get_state(self._h, *[_p(out[k]) for k in ("qpos", "qvel", "target", "fingertip", "step", "episode", "qpos_lo")])
        return out

    def set_state(self, qpos=None, qvel=None, target=None, fingertip=None, step=None, episode=None, qpos_lo=None):
        def prep(a, dt):
            return None if a is None else np.ascontiguousarray(a, dt)
        arrs = [prep(qpos, np.float32), prep(qvel, np.float32), prep(target, np.float32), prep(fingertip, np.float32), prep(step, np.int32),
                prep(episode, np.uint32), prep(qpos_lo, np.float32)]
        lib().twin_set_state(self._h, *[_p(a) for a in arrs])
