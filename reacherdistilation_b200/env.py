"""Gym-style surface over the device-resident lock-step Reacher-v2 environments.

Mirrors what the reference uses of its env object (/root/reference src/distilation/mlp_train.py:21,33,112,135; teacher.py:15,37):
reset() -> ob, step(a) -> (ob, reward, done, info), observation_space.shape == (11,), action_space a 2-d Box in [-1,1], close().
Batched generalisation: leading dim N on every tensor, done[N], finished envs auto-reset and return their reset observation.
Returned tensors are views of buffers owned by the env, valid until the next step()/reset().
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class Box:
    def __init__(self, low, high, shape):
        self.low = np.full(shape, low, dtype=np.float32)
        self.high = np.full(shape, high, dtype=np.float32)
        self.shape = tuple(shape)
        self.dtype = np.float32


class VecReacher:
    """N Reacher-v2 environments on one B200.  host=True gives the numpy-in / numpy-out surface (copies inside the call)."""

    def __init__(self, num_envs=1, seed=0, device=0, env_offset=0, host=False, squeeze=False):
        import ctypes as C
        if not torch.cuda.is_available():
            raise _lib.ReacherB200Error("reacherdistilation_b200 needs a CUDA device (there is no CPU path)")
        self.num_envs, self.seed, self.env_offset = int(num_envs), int(seed), int(env_offset)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.host, self.squeeze = host, squeeze and num_envs == 1
        self.observation_space = Box(-np.inf, np.inf, (11,))
        self.action_space = Box(-1.0, 1.0, (2,))
        h = C.c_void_p()
        check(lib().rb_env_create(C.byref(h), self.num_envs, self.seed, self.device.index, self.env_offset))
        self._h = h
        n = self.num_envs
        if host:
            self._obs = torch.empty((n, 11), dtype=torch.float32).pin_memory()
            self._rew = torch.empty((n,), dtype=torch.float32).pin_memory()
            self._done = torch.empty((n,), dtype=torch.uint8).pin_memory()
            self._act = torch.empty((n, 2), dtype=torch.float32).pin_memory()
            self._pd = torch.empty((n, 4), dtype=torch.float32).pin_memory()
            # numpy views + raw addresses, taken once: the per-step path below is a few hundred nanoseconds of Python around one C call
            self._obs_np, self._rew_np, self._done_np, self._act_np, self._pd_np = (t.numpy() for t in (self._obs, self._rew, self._done, self._act, self._pd))
            self._p_obs, self._p_rew, self._p_done, self._p_act, self._p_pd = (t.data_ptr() for t in (self._obs, self._rew, self._done, self._act, self._pd))
            self._policy = None            # TeacherAgent registered with the resident env server (N <= 32): step() also returns its pdflat
            self._last_ob = self._last_pd = None
        else:
            with torch.cuda.device(self.device):
                self._obs = torch.empty((n, 11), dtype=torch.float32, device=self.device)
                self._rew = torch.empty((n,), dtype=torch.float32, device=self.device)
                self._done = torch.empty((n,), dtype=torch.uint8, device=self.device)

    SERVE_MAX_ENVS = 32                    # csrc/serve.cu

    def serve_policy(self, agent, params_host, nout):
        """Register a policy with the resident env server (host surface, N <= 32): every step() then also evaluates it on the observation it
        returns (one round trip for `ac = pi.act(ob); ob = env.step(ac)`, mlp_train.py:123-135).  Returns False when not applicable."""
        if not self.host or self.num_envs > self.SERVE_MAX_ENVS:
            return False
        self._params_keep = np.ascontiguousarray(params_host, np.float32)
        check(lib().rb_env_serve_policy(self._h, self._params_keep.ctypes.data, int(nout)))
        self._policy, self._last_ob, self._last_pd = agent, None, None
        return True

    def _host_ob(self, with_pd):
        ob = self._obs_np.copy()                                       # a fresh array per call, like gym (callers keep observations in lists)
        ob = ob[0] if self.squeeze else ob
        if with_pd:
            self._last_ob, self._last_pd = ob, self._pd_np.copy()
        return ob

    # ---- gym protocol --------------------------------------------------------------------------------------
    def reset(self):
        if self.host:
            check(lib().rb_env_reset_host(self._h, self._p_obs))
            self._last_ob = self._last_pd = None
            return self._host_ob(False)
        check(lib().rb_env_reset(self._h, ptr(self._obs), stream_ptr()))
        return self._obs

    def step(self, action):
        n = self.num_envs
        if self.host:
            if torch.is_tensor(action):
                action = action.detach().cpu().numpy()
            self._act_np[...] = np.asarray(action, dtype=np.float32).reshape(n, 2)
            served = self._policy is not None
            if served:
                rc = lib().rb_env_act_step_host(self._h, self._p_act, self._p_obs, self._p_rew, self._p_done, self._p_pd)
            else:
                rc = lib().rb_env_step_host(self._h, self._p_act, self._p_obs, self._p_rew, self._p_done)
            if rc:
                check(rc)
            ob = self._host_ob(served)
            if self.squeeze:
                return ob, float(self._rew_np[0]), bool(self._done_np[0]), {}
            return ob, self._rew_np.copy(), self._done_np.astype(bool), {}
        a = action
        assert a.is_cuda and a.dtype == torch.float32 and a.numel() == 2 * n, "action must be a float32 CUDA tensor [N,2]"
        a = a.contiguous()
        check(lib().rb_env_step(self._h, ptr(a), ptr(self._obs), ptr(self._rew), ptr(self._done), stream_ptr()))
        return self._obs, self._rew, self._done, {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().rb_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed_(self, seed):
        raise NotImplementedError("seed is fixed at construction (Philox key)")

    # ---- device-side extras --------------------------------------------------------------------------------
    def observe(self):
        assert not self.host
        check(lib().rb_env_observe(self._h, ptr(self._obs), stream_ptr()))
        return self._obs

    def get_state(self):
        n, dev = self.num_envs, self.device
        out = dict(qpos=torch.empty((n, 2), device=dev), qvel=torch.empty((n, 2), device=dev), target=torch.empty((n, 2), device=dev),
                   fingertip=torch.empty((n, 2), device=dev), step=torch.empty((n,), dtype=torch.int32, device=dev),
                   episode=torch.empty((n,), dtype=torch.int32, device=dev), qpos_lo=torch.empty((n, 2), device=dev))
        check(lib().rb_env_get_state(self._h, ptr(out["qpos"]), ptr(out["qvel"]), ptr(out["target"]), ptr(out["fingertip"]),
                                     ptr(out["step"]), ptr(out["episode"]), ptr(out["qpos_lo"]), stream_ptr()))
        return out

    def set_state(self, qpos=None, qvel=None, target=None, fingertip=None, step=None, episode=None, qpos_lo=None):
        """qpos_lo: low parts of the two-float joint angles (get_state()["qpos_lo"]); omitted -> zero (qpos is then the whole angle)."""
        def prep(t, dt):
            return None if t is None else torch.as_tensor(t, dtype=dt).to(self.device).contiguous()
        ts = [prep(qpos, torch.float32), prep(qvel, torch.float32), prep(target, torch.float32), prep(fingertip, torch.float32),
              prep(step, torch.int32), prep(episode, torch.int32), prep(qpos_lo, torch.float32)]
        check(lib().rb_env_set_state(self._h, *[ptr(t) for t in ts], stream_ptr()))
        torch.cuda.current_stream().synchronize()   # keep the temporaries alive until the kernel has read them

    def rollout_random(self, T, step0=0, record_obs=True, record_act=False):
        """T fused steps with Philox U(-1,1) actions.  Returns dict of time-major device buffers."""
        n, dev = self.num_envs, self.device
        out = dict(obs=torch.empty((T, n, 11), device=dev) if record_obs else None,
                   act=torch.empty((T, n, 2), device=dev) if record_act else None,
                   rew=torch.empty((T, n), device=dev), done=torch.empty((T, n), dtype=torch.uint8, device=dev))
        check(lib().rb_env_rollout_random(self._h, T, step0, ptr(out["obs"]), ptr(out["act"]), ptr(out["rew"]), ptr(out["done"]), stream_ptr()))
        return out

    def rollout_policy(self, params, T, nout=2, mode=_lib.MODE_FP32, out=None):
        """Fused policy-in-the-loop rollout into a device-resident buffer (teacher warm-up, mlp_train.py:120-139)."""
        n, dev = self.num_envs, self.device
        if out is None:
            out = dict(obs=torch.empty((T, n, 11), device=dev), pdflat=torch.empty((T, n, 4), device=dev),
                       rew=torch.empty((T, n), device=dev), done=torch.empty((T, n), dtype=torch.uint8, device=dev))
        check(lib().rb_env_rollout_policy(self._h, ptr(params), nout, T, ptr(out["obs"]), ptr(out["pdflat"]), ptr(out["rew"]),
                                          ptr(out["done"]), mode, stream_ptr()))
        return out

    def rollout_policy_host(self, params_host, T, nout=2, mode=_lib.MODE_FP32, out=None):
        """Same through HOST buffers (H2D of the parameters, D2H inside the call).  The rollout buffer is always written on the device
        (rollout_buffer()); `out` entries that are None are not brought to the host."""
        n = self.num_envs
        if out is None:
            out = dict(obs=torch.empty((T, n, 11)).pin_memory(), pdflat=torch.empty((T, n, 4)).pin_memory(),
                       rew=torch.empty((T, n)).pin_memory(), done=torch.empty((T, n), dtype=torch.uint8).pin_memory())
        if out.get("done_mask") is not None or out.get("return_sum") is not None:
            # done_mask [N] int64: bit t = the env finished an episode at step t of this call (T <= 64); return_sum [N] float32: sum of the call's rewards
            assert out.get("done_mask") is None or (out["done_mask"].dtype == torch.int64 and out["done_mask"].numel() == n)
            assert out.get("return_sum") is None or (out["return_sum"].dtype == torch.float32 and out["return_sum"].numel() == n)
            check(lib().rb_env_rollout_policy_host_ex(self._h, ptr(params_host), nout, T, ptr(out.get("obs")), ptr(out.get("pdflat")), ptr(out.get("rew")),
                                                      ptr(out.get("done")), ptr(out.get("done_mask")), ptr(out.get("return_sum")), mode))
            return out
        check(lib().rb_env_rollout_policy_host(self._h, ptr(params_host), nout, T, ptr(out.get("obs")), ptr(out.get("pdflat")), ptr(out.get("rew")),
                                               ptr(out.get("done")), mode))                  # None entries: that field is not copied to the host
        return out

    def rollout_policy_host_begin(self, params_host, T, nout=2, mode=_lib.MODE_FP32, out=None):
        """Split-phase host rollout: queue {H2D of the parameters, the launch} and return; rollout_policy_host_wait() completes the OLDEST
        outstanding call.  Up to two in flight.  out: dict of page-locked tensors rew [T,N] f32 / done_mask [N] int64 / return_sum [N] f32 (any
        may be absent); they must not be touched until the matching wait."""
        check(lib().rb_env_rollout_policy_host_begin(self._h, ptr(params_host), nout, T, ptr(out.get("rew")), ptr(out.get("done_mask")),
                                                     ptr(out.get("return_sum")), mode))
        return out

    def rollout_policy_host_wait(self):
        check(lib().rb_env_rollout_policy_host_wait(self._h))

    def rollout_buffer(self):
        """Device views (no copy) of the resident rollout buffer the last rollout_policy_host() filled: dict(obs [T,N,11], pdflat [T,N,4],
        rew [T,N], done [T,N]).  rew / done only hold data when that call's host buffers were pageable or absent (page-locked ones are
        written by the kernel directly).  Valid until the next host rollout or close()."""
        import ctypes as C
        po, pp, pr, pdn, T = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        check(lib().rb_env_rollout_buffer(self._h, C.byref(po), C.byref(pp), C.byref(pr), C.byref(pdn), C.byref(T)))
        n, T = self.num_envs, T.value

        class _View:
            def __init__(self, addr, shape, typestr):
                self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(addr, False), version=2, strides=None)
        with torch.cuda.device(self.device):
            return dict(obs=torch.as_tensor(_View(po.value, (T, n, 11), "<f4"), device=self.device),
                        pdflat=torch.as_tensor(_View(pp.value, (T, n, 4), "<f4"), device=self.device),
                        rew=torch.as_tensor(_View(pr.value, (T, n), "<f4"), device=self.device),
                        done=torch.as_tensor(_View(pdn.value, (T, n), "|u1"), device=self.device))


def make_mujoco_env(env_id, seed, num_envs=1, device=0, rank=0, host=None):
    """baselines.common.cmd_util.make_mujoco_env as the reference calls it (mlp_train.py:21): seed + 1000 * rank.
    num_envs == 1 returns the host (numpy) surface of a single env, like gym; larger N returns device tensors."""
    if env_id != "Reacher-v2":
        raise ValueError("only Reacher-v2 is implemented (the reference uses nothing else)")
    host = (num_envs == 1) if host is None else host
    return VecReacher(num_envs=num_envs, seed=seed + 1000 * rank, device=device, host=host, squeeze=host)
