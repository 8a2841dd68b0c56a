"""Teacher policy kernels vs the oracle: standalone forward and the fused policy-in-the-loop rollout (BASELINE config 3)."""
import numpy as np
import pytest
import torch

from oracle import nn_np as NN
from oracle import reacher_c as RC

pytestmark = pytest.mark.gpu


def _modes():
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200._lib import lib
    out = [("fp32", MODE_FP32, 1e-6)]              # measured 2.4e-7 (fp32 FMA order vs float64) / 6.6e-6 (tcgen05, bf16x3 operands): gates at ~3x
    if lib().rb_mode_available(MODE_TC):
        out.append(("tc", MODE_TC, 2e-5))
    return out


@pytest.mark.parametrize("nout", [2, 4])
@pytest.mark.parametrize("n", [1, 127, 128, 4097])
def test_policy_forward_matches_oracle(n, nout):
    from reacherdistilation_b200._lib import check, lib, ptr, stream_ptr
    from reacherdistilation_b200.teacher import init_policy_params
    rng = np.random.default_rng(n + nout)
    p = init_policy_params(seed=3, nout=nout, final_std=0.5, ob_mean=rng.standard_normal(11) * 0.1, ob_std=1 + rng.random(11))
    obs = (rng.standard_normal((n, 11)) * 3).astype(np.float32)     # exercises the +-5 clip
    ref = NN.policy_fwd(obs, p, nout=nout)
    p_dev, obs_dev = torch.from_numpy(p).cuda(), torch.from_numpy(obs).cuda()      # keep alive across the async launch
    for name, mode, tol in _modes():
        out = torch.empty((n, 4), device="cuda")
        check(lib().rb_policy_fwd(ptr(p_dev), nout, ptr(obs_dev), n, ptr(out), mode, stream_ptr()))
        err = np.abs(out.cpu().numpy() - ref).max()
        print("policy_fwd %s n=%d nout=%d max err %.3g" % (name, n, nout, err))
        assert err <= tol, (name, err)


def test_policy_forward_host_entry_point():
    from reacherdistilation_b200._lib import MODE_FP32, check, lib
    from reacherdistilation_b200.teacher import init_policy_params
    p = init_policy_params(seed=1)
    obs = np.random.default_rng(0).standard_normal((100, 11)).astype(np.float32)
    out = np.zeros((100, 4), np.float32)
    check(lib().rb_policy_fwd_host(p.ctypes.data, 2, obs.ctypes.data, 100, out.ctypes.data, MODE_FP32, 0))
    assert np.abs(out - NN.policy_fwd(obs, p)).max() <= 2e-6
    assert np.allclose(out[:, 2:], p[-2:])


@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
def test_fused_teacher_rollout_teacher_forced_policy_and_action_forced_physics(mode_name):
    """THE parity gate of the fused rollout (60 steps incl. an auto-reset), free of closed-loop amplification:
      * teacher-forced: the recorded pdflat is the oracle policy of the recorded observation, to kernel precision (1e-6 fp32, 2e-5 tcgen05);
      * action-forced: the recorded observations / done flags are, BIT FOR BIT, what the host build of csrc/physics.cuh (tests/twin/) produces
        when it is stepped with the recorded actions -- the physics inside the fused kernel is exactly the arithmetic whose tolerance against
        the float64 oracle tests/test_physics_twin_cpu.py states (rewards: to the last ulp of the MUFU square root)."""
    from twin.physics_twin import PhysicsTwin
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    sel = [m for m in _modes() if m[0] == mode_name]
    if not sel:
        pytest.skip("RB_MODE_TC not compiled into this build")
    name, mode, ptol = sel[0]
    n, T, seed = 1000, 60, 2
    p = init_policy_params(seed=0, final_std=0.3)
    env = VecReacher(num_envs=n, seed=seed)
    ob0 = env.reset().cpu().numpy()
    out = env.rollout_policy(torch.from_numpy(p).cuda(), T, nout=2, mode=mode)
    d_obs, d_pd, d_rew, d_done = (out[k].cpu().numpy() for k in ("obs", "pdflat", "rew", "done"))
    ref_pd = NN.policy_fwd(d_obs.reshape(-1, 11), p).reshape(T, n, 4)
    e_pd = np.abs(d_pd - ref_pd).max()
    tw = PhysicsTwin(n, seed=seed)
    assert np.array_equal(tw.reset(), ob0) and np.array_equal(d_obs[0], ob0)
    e_rew = 0.0
    for t in range(T):
        ob, rw, dn = tw.step(d_pd[t, :, :2])
        if t + 1 < T:
            assert np.array_equal(ob, d_obs[t + 1]), "fused-kernel physics differs from the host build at step %d" % t
        assert np.array_equal(dn, d_done[t].astype(bool))
        e_rew = max(e_rew, float(np.abs(rw - d_rew[t]).max()))
    print("fused rollout %s: teacher-forced pdflat err %.3g (gate %.1g); action-forced obs bit-identical to the host twin, reward err %.3g" % (name, e_pd, ptol, e_rew))
    assert e_pd <= ptol and e_rew <= 5e-7
    env.close()


@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
def test_fused_teacher_rollout_closed_loop_distribution_vs_oracle(mode_name):
    """Closed loop (policy in the loop) vs the C oracle running the same policy: informative, not the parity gate.  A difference d in the
    action feeds back through gear 200 and the stiff joint limit (an oracle-vs-oracle perturbation study gives 1e-5 -> 1e-1 on the worst of
    ~1000 envs within an episode), so the gate is on the distribution over envs of the per-env worst error, at ~4x the measured values:
    median / 99 % / max  fp32 1.6e-6 / 4.1e-5 / 3.9e-4 -> 1e-5 / 2e-4 / 2e-3,  tcgen05 2.5e-5 / 5.7e-4 / 2.4e-2 -> 1e-4 / 2.5e-3 / 1e-1."""
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    sel = [m for m in _modes() if m[0] == mode_name]
    if not sel:
        pytest.skip("RB_MODE_TC not compiled into this build")
    name, mode, ptol = sel[0]
    n, T, seed = 1000, 60, 2
    p = init_policy_params(seed=0, final_std=0.3)
    env = VecReacher(num_envs=n, seed=seed)
    env.reset()
    out = env.rollout_policy(torch.from_numpy(p).cuda(), T, nout=2, mode=mode)
    c = RC.ReacherOracleC(n, seed=seed); c.reset()
    obs, pd, rew, done, _ = c.rollout_policy(T, p)
    assert np.array_equal(out["done"].cpu().numpy(), done)
    d_obs, d_pd, d_rew = out["obs"].cpu().numpy(), out["pdflat"].cpu().numpy(), out["rew"].cpu().numpy()
    e_env = np.maximum(np.maximum(np.abs(d_obs - obs).max(axis=(0, 2)), np.abs(d_pd - pd).max(axis=(0, 2))), np.abs(d_rew - rew).max(axis=0))
    print("fused rollout %s closed loop, per-env worst error: median %.3g, 99%% %.3g, max %.3g" % (name, np.median(e_env), np.quantile(e_env, 0.99), e_env.max()))
    lim = (1e-5, 2e-4, 2e-3) if name == "fp32" else (1e-4, 2.5e-3, 1e-1)
    assert np.median(e_env) <= lim[0] and np.quantile(e_env, 0.99) <= lim[1] and e_env.max() <= lim[2], (np.median(e_env), np.quantile(e_env, 0.99), e_env.max())
    env.close()


def test_rollout_host_entry_point_matches_device():
    from reacherdistilation_b200 import MODE_FP32
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    p = init_policy_params(seed=0)
    a, b = VecReacher(num_envs=300, seed=1), VecReacher(num_envs=300, seed=1)
    a.reset(); b.reset()
    d = a.rollout_policy(torch.from_numpy(p).cuda(), 55, mode=MODE_FP32)
    h = b.rollout_policy_host(torch.from_numpy(p), 55, mode=MODE_FP32)
    for k in d:
        assert torch.equal(d[k].cpu(), h[k]), k
    a.close(); b.close()


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
def test_rollout_host_result_only_keeps_buffer_on_device(pinned, mode_name):
    """reward / done to the host (kernel-written when the host buffers are page-locked, slab copies when pageable), obs / pdflat left in
    the device-resident buffer: every field bit-identical to the device entry point on the same seed."""
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    mode = MODE_FP32 if mode_name == "fp32" else MODE_TC
    p = init_policy_params(seed=0)
    n, T = 333, 53                                        # ragged warp, slabs of unequal length, one auto-reset inside
    a, b = VecReacher(num_envs=n, seed=4), VecReacher(num_envs=n, seed=4)
    a.reset(); b.reset()
    d = a.rollout_policy(torch.from_numpy(p).cuda(), T, mode=mode)
    rew, done = torch.full((T, n), -7.0), torch.full((T, n), 9, dtype=torch.uint8)
    if pinned:
        rew, done = rew.pin_memory(), done.pin_memory()
    h = b.rollout_policy_host(torch.from_numpy(p), T, mode=mode, out=dict(obs=None, pdflat=None, rew=rew, done=done))
    assert torch.equal(d["rew"].cpu(), h["rew"]) and torch.equal(d["done"].cpu(), h["done"])
    buf = b.rollout_buffer()
    assert buf["obs"].shape == (T, n, 11) and buf["pdflat"].shape == (T, n, 4)
    assert torch.equal(buf["obs"], d["obs"]) and torch.equal(buf["pdflat"], d["pdflat"])
    if not pinned:
        assert torch.equal(buf["rew"], d["rew"]) and torch.equal(buf["done"], d["done"])
    # the env state advanced identically: one more chunk through the other entry point still agrees
    d2 = a.rollout_policy(torch.from_numpy(p).cuda(), 7, mode=mode)
    h2 = b.rollout_policy_host(torch.from_numpy(p), 7, mode=mode)
    for k in d2:
        assert torch.equal(d2[k].cpu(), h2[k]), k
    a.close(); b.close()


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
def test_rollout_host_done_mask_is_the_bit_packed_done(pinned, mode_name):
    """rb_env_rollout_policy_host_ex: done_mask[N] (bit t = the env finished an episode at step t of the call) == the u8 done[T,N] of the
    device entry point, bit for bit, with reward bit-identical and no u8 `done` transferred at all; also next to a u8 `done` in the same call."""
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    mode = MODE_FP32 if mode_name == "fp32" else MODE_TC
    p = init_policy_params(seed=0)
    n, T = 333, 64
    a, b = VecReacher(num_envs=n, seed=4), VecReacher(num_envs=n, seed=4)
    a.reset(); b.reset()
    a.rollout_policy(torch.from_numpy(p).cuda(), 17, mode=mode); b.rollout_policy(torch.from_numpy(p).cuda(), 17, mode=mode)   # episode boundary not at a multiple of the call
    for with_u8 in (False, True):
        d = a.rollout_policy(torch.from_numpy(p).cuda(), T, mode=mode)
        rew, mask, ret = torch.full((T, n), -7.0), torch.full((n,), -1, dtype=torch.int64), torch.full((n,), 3.0)
        done = torch.full((T, n), 9, dtype=torch.uint8) if with_u8 else None
        if pinned:
            rew, mask, ret = rew.pin_memory(), mask.pin_memory(), ret.pin_memory()
            done = done.pin_memory() if done is not None else None
        b.rollout_policy_host(torch.from_numpy(p), T, mode=mode, out=dict(obs=None, pdflat=None, rew=rew, done=done, done_mask=mask, return_sum=ret))
        acc = np.zeros(n, np.float32)
        for t in range(T):                                     # the kernel adds in step order, in fp32: bit-exact
            acc = (acc + d["rew"][t].cpu().numpy()).astype(np.float32)
        assert np.array_equal(ret.numpy(), acc)
        want = np.zeros(n, np.uint64)
        dd = d["done"].cpu().numpy()
        for t in range(T):
            want |= dd[t].astype(np.uint64) << np.uint64(t)
        assert np.array_equal(mask.numpy().view(np.uint64), want) and dd.sum() > 0
        assert torch.equal(d["rew"].cpu(), rew) and (done is None or torch.equal(d["done"].cpu(), done))
    with pytest.raises(Exception):
        b.rollout_policy_host(torch.from_numpy(p), 65, mode=mode, out=dict(obs=None, pdflat=None, rew=None, done=None, done_mask=mask))
    a.close(); b.close()


@pytest.mark.parametrize("transport", [1, 0])
@pytest.mark.parametrize("mode_name", ["fp32", "tc"])
def test_split_phase_host_rollout_two_in_flight(mode_name, transport):
    """rb_env_rollout_policy_host_begin / _wait: two calls in flight, outputs of call i complete after its wait while call i + 1 runs; the
    trajectories equal the synchronous entry point's bit for bit; a third begin without a wait is refused."""
    from reacherdistilation_b200 import MODE_FP32, MODE_TC
    from reacherdistilation_b200._lib import ReacherB200Error
    from reacherdistilation_b200.env import VecReacher
    from reacherdistilation_b200.teacher import init_policy_params
    mode = MODE_FP32 if mode_name == "fp32" else MODE_TC
    p = torch.from_numpy(init_policy_params(seed=0)).pin_memory()
    n, T, calls = 4099, 50, 5
    a, b = VecReacher(num_envs=n, seed=6), VecReacher(num_envs=n, seed=6)
    a.reset(); b.reset()
    from reacherdistilation_b200 import _lib
    _lib.check(_lib.lib().rb_env_set_host_transport(b._h, transport))        # 1: reward stored by the kernel; 0: copy engine beside the next kernel
    want = [a.rollout_policy(p.cuda(), T, mode=mode) for _ in range(calls)]
    want = [{k: v.clone() for k, v in w.items()} for w in want]
    bufs = [dict(rew=torch.full((T, n), -7.0).pin_memory(), done_mask=torch.full((n,), -1, dtype=torch.int64).pin_memory(),
                 return_sum=torch.full((n,), 3.0).pin_memory()) for _ in range(2)]
    got = []
    for i in range(calls):
        b.rollout_policy_host_begin(p, T, mode=mode, out=bufs[i & 1])
        if i == 1:
            with pytest.raises(ReacherB200Error):
                b.rollout_policy_host_begin(p, T, mode=mode, out=bufs[0])          # two are in flight
        if i >= 1:
            b.rollout_policy_host_wait()
            got.append({k: v.clone() for k, v in bufs[(i - 1) & 1].items()})
    b.rollout_policy_host_wait()
    got.append({k: v.clone() for k, v in bufs[(calls - 1) & 1].items()})
    with pytest.raises(ReacherB200Error):
        b.rollout_policy_host_wait()
    for w, g in zip(want, got):
        assert torch.equal(w["rew"].cpu(), g["rew"])
        dd = w["done"].cpu().numpy()
        m = np.zeros(n, np.uint64)
        for t in range(T):
            m |= dd[t].astype(np.uint64) << np.uint64(t)
        assert np.array_equal(g["done_mask"].numpy().view(np.uint64), m)
        acc = np.zeros(n, np.float32)
        for t in range(T):
            acc = (acc + w["rew"][t].cpu().numpy()).astype(np.float32)
        assert np.array_equal(g["return_sum"].numpy(), acc)
    a.close(); b.close()
