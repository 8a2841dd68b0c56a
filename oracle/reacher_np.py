"""ORACLE (test infrastructure, not product code) -- float64 numpy restatement of Reacher-v2.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

What it restates.  The reference never implements the env itself: it calls the third-party
gym==0.10.5 ReacherEnv on MuJoCo 1.50 via mujoco-py==1.50.1.56 (pins: /root/reference
src/distilation/requirement.txt:20,33) at these call sites:
  * env creation      src/distilation/mlp_train.py:21, lstm_train.py:21   make_mujoco_env("Reacher-v2", 0)
  * env.reset()       src/distilation/mlp_train.py:112,138,200
  * env.step(a)       src/distilation/mlp_train.py:135,196 ; lstm_train.py:133,192
None of gym / MuJoCo is present in /root/reference, so the published model (gym reacher.xml + MuJoCo RK4
semantics) is restated here and PINNED against the reference's own recorded MuJoCo trajectories,
src/distilation/tests/data/dataset.json (committed as tests/golden/reacher_fixture.npz by
tests/golden/make_golden.py): all 25x49 one-step transitions reproduce to < 1e-13 (tests/test_oracle_fixture.py).

Parity status: env step / obs / reward / episode length / joint limit / stale kinematics = PINNED by fixture.
Reset RNG stream (gym's MT19937) = replaced by Philox4x32-10 by design (north_star) -> only the reset RANGES are
pinned by the fixture; the stream itself is defined here and must be bit-exact between this file, reacher_oracle.c
and the CUDA kernel.
"""
import numpy as np

from .philox_np import philox4x32_10, u32_to_uniform_f32

# ---- model constants (gym reacher.xml under MuJoCo-1.50 capsule-inertia rules) ------------
RHO, R_CAP, L_LINK, L0, L_TIP = 1000.0, 0.01, 0.1, 0.1, 0.11
M_LINK = RHO * np.pi * R_CAP ** 2 * (L_LINK + R_CAP)            # capsule == cylinder of length L+r in MuJoCo 1.50
I_LINK = M_LINK * (3 * R_CAP ** 2 + (L_LINK + R_CAP) ** 2) / 12.0
M_TIP = RHO * (4.0 / 3.0) * np.pi * R_CAP ** 3
I_TIP = 0.4 * M_TIP * R_CAP ** 2
C_ = I_LINK + M_LINK * 0.05 ** 2 + I_TIP + M_TIP * L_TIP ** 2
B_ = L0 * (M_LINK * 0.05 + M_TIP * L_TIP)
A_ = I_LINK + M_LINK * 0.05 ** 2 + C_ + (M_LINK + M_TIP) * L0 ** 2
ARMATURE, DAMPING, GEAR = 1.0, 1.0, 200.0
H, FRAME_SKIP, EPISODE_STEPS = 0.01, 2, 50
LIMIT = 3.0
SOLREF_TC, SOLREF_DR = 0.02, 1.0
SOLIMP_DMIN, SOLIMP_DMAX, SOLIMP_W = 0.9, 0.95, 0.001
K_LIM = 1.0 / (SOLIMP_DMAX ** 2 * SOLREF_TC ** 2 * SOLREF_DR ** 2)
B_LIM = 2.0 / (SOLIMP_DMAX * SOLREF_TC)
# dof_invweight0 of joint1: (M(theta1=0)^-1)[1,1]
_M00_0, _M01_0, _M11_0 = ARMATURE + A_ + 2 * B_, C_ + B_, ARMATURE + C_
INVW0 = _M00_0 / (_M00_0 * _M11_0 - _M01_0 ** 2)

RESET_QPOS, RESET_GOAL, RESET_QVEL = 0.1, 0.2, 0.005
STREAM_RESET, STREAM_ACTION, STREAM_DROPOUT = 0, 1, 2


def accel(q1, v0, v1, u0, u1):
    """Joint accelerations (mj_forward restated).  u already clipped to ctrlrange."""
    c1, s1 = np.cos(q1), np.sin(q1)
    m00 = ARMATURE + A_ + 2 * B_ * c1
    m01 = C_ + B_ * c1
    m11 = ARMATURE + C_
    cor0 = -B_ * s1 * (2 * v0 * v1 + v1 * v1)
    cor1 = B_ * s1 * v0 * v0
    t0 = GEAR * u0 - DAMPING * v0 - cor0
    t1 = GEAR * u1 - DAMPING * v1 - cor1
    det = m00 * m11 - m01 * m01
    a0 = (m11 * t0 - m01 * t1) / det
    a1 = (m00 * t1 - m01 * t0) / det
    mi01, mi11 = -m01 / det, m00 / det
    # soft joint limit on theta1 (range +-3, margin 0)
    for sgn, dist in ((-1.0, LIMIT - q1), (+1.0, q1 + LIMIT)):
        act = dist < 0
        if not np.any(act):
            continue
        x = np.minimum(1.0, np.abs(dist) / SOLIMP_W)
        y = np.where(x <= 0.5, 2 * x * x, 1 - 2 * (1 - x) ** 2)
        imp = SOLIMP_DMIN + (SOLIMP_DMAX - SOLIMP_DMIN) * y
        aref = -B_LIM * (sgn * v1) - K_LIM * imp * dist
        Rr = (1 - imp) / imp * INVW0
        f = np.maximum(0.0, (aref - sgn * a1) / (mi11 + Rr))
        f = np.where(act, f, 0.0)
        a0 = a0 + mi01 * sgn * f
        a1 = a1 + mi11 * sgn * f
    return a0, a1


def fk(q0, q1):
    return (L0 * np.cos(q0) + L_TIP * np.cos(q0 + q1), L0 * np.sin(q0) + L_TIP * np.sin(q0 + q1))


def rk4_substep(q0, q1, v0, v1, u0, u1):
    """One mj_step with mjINT_RK4; returns new state and the LAST-STAGE qpos (MuJoCo's stale kinematics)."""
    h = H
    f00, f01 = accel(q1, v0, v1, u0, u1)
    qa0, qa1, va0, va1 = q0 + h / 2 * v0, q1 + h / 2 * v1, v0 + h / 2 * f00, v1 + h / 2 * f01
    f10, f11 = accel(qa1, va0, va1, u0, u1)
    qb0, qb1, vb0, vb1 = q0 + h / 2 * va0, q1 + h / 2 * va1, v0 + h / 2 * f10, v1 + h / 2 * f11
    f20, f21 = accel(qb1, vb0, vb1, u0, u1)
    qc0, qc1, vc0, vc1 = q0 + h * vb0, q1 + h * vb1, v0 + h * f20, v1 + h * f21
    f30, f31 = accel(qc1, vc0, vc1, u0, u1)
    nq0 = q0 + h / 6 * (v0 + 2 * va0 + 2 * vb0 + vc0)
    nq1 = q1 + h / 6 * (v1 + 2 * va1 + 2 * vb1 + vc1)
    nv0 = v0 + h / 6 * (f00 + 2 * f10 + 2 * f20 + f30)
    nv1 = v1 + h / 6 * (f01 + 2 * f11 + 2 * f21 + f31)
    return nq0, nq1, nv0, nv1, qc0, qc1


def make_obs(q0, q1, v0, v1, tx, ty, px, py):
    z = np.zeros_like(q0)
    return np.stack([np.cos(q0), np.cos(q1), np.sin(q0), np.sin(q1), tx, ty, v0, v1, px - tx, py - ty, z], -1)


def reset_draws(seed, env_ids, episodes):
    """Philox reset sampling, float32-exact (values are float32 numbers, returned as float64).

    counter = (env_id, episode_idx, draw_idx, STREAM_RESET), key = (seed_lo, seed_hi).
    draw 0 -> (qpos0, qpos1, goal_x, goal_y), draw 1 -> (qvel0, qvel1, -, -).
    gym's reset_model (third-party; call site mlp_train.py:112) draws qpos noise U(+-0.1), goal U(+-0.2)^2 with a
    ||g||<2 rejection that can never fire, qvel noise U(+-0.005); ranges pinned by the fixture.
    """
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    episodes = np.asarray(episodes, dtype=np.uint32)
    r0 = philox4x32_10(seed, env_ids, episodes, np.zeros_like(env_ids), np.full_like(env_ids, STREAM_RESET))
    r1 = philox4x32_10(seed, env_ids, episodes, np.ones_like(env_ids), np.full_like(env_ids, STREAM_RESET))
    f = np.float32
    q0 = u32_to_uniform_f32(r0[0], f(-RESET_QPOS), f(RESET_QPOS))
    q1 = u32_to_uniform_f32(r0[1], f(-RESET_QPOS), f(RESET_QPOS))
    tx = u32_to_uniform_f32(r0[2], f(-RESET_GOAL), f(RESET_GOAL))
    ty = u32_to_uniform_f32(r0[3], f(-RESET_GOAL), f(RESET_GOAL))
    v0 = u32_to_uniform_f32(r1[0], f(-RESET_QVEL), f(RESET_QVEL))
    v1 = u32_to_uniform_f32(r1[1], f(-RESET_QVEL), f(RESET_QVEL))
    return tuple(a.astype(np.float64) for a in (q0, q1, v0, v1, tx, ty))


def random_actions(seed, env_ids, step_idx):
    """a ~ U(-1,1)^2 keyed (seed, env, global step) -- SURVEY 8(d) config 2 synthetic action stream."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    st = np.full_like(env_ids, np.uint32(step_idx))
    r = philox4x32_10(seed, env_ids, st, np.zeros_like(env_ids), np.full_like(env_ids, STREAM_ACTION))
    a0 = u32_to_uniform_f32(r[0], np.float32(-1), np.float32(1))
    a1 = u32_to_uniform_f32(r[1], np.float32(-1), np.float32(1))
    return np.stack([a0, a1], -1)  # float32


class ReacherOracle:
    """Vectorised float64 Reacher-v2 with TimeLimit(50) and auto-reset (gym Monitor/TimeLimit restated).

    state_dtype=np.float32 rounds the carried state after every env step (models an fp32-state device kernel
    with exact arithmetic) -- used only to derive the stated tolerance.
    """

    def __init__(self, num_envs, seed=0, env_offset=0):
        self.n = int(num_envs)
        self.seed = int(seed)
        self.env_ids = (np.arange(self.n, dtype=np.uint64) + env_offset).astype(np.uint32)
        self.episode = np.zeros(self.n, dtype=np.uint32)
        self.step_count = np.zeros(self.n, dtype=np.int32)
        self.q0 = self.q1 = self.v0 = self.v1 = self.tx = self.ty = self.px = self.py = None

    # -- gym surface ------------------------------------------------------------------------------------------
    def reset(self):
        self.episode[:] = 0
        self._reset_where(np.ones(self.n, bool))
        return self.obs()

    def _reset_where(self, mask):
        q0, q1, v0, v1, tx, ty = reset_draws(self.seed, self.env_ids, self.episode)
        px, py = fk(q0, q1)
        if self.q0 is None:
            self.q0, self.q1, self.v0, self.v1, self.tx, self.ty, self.px, self.py = q0, q1, v0, v1, tx, ty, px, py
        else:
            for name, val in (("q0", q0), ("q1", q1), ("v0", v0), ("v1", v1), ("tx", tx), ("ty", ty), ("px", px), ("py", py)):
                setattr(self, name, np.where(mask, val, getattr(self, name)))
        self.step_count = np.where(mask, 0, self.step_count).astype(np.int32)

    def set_state(self, q0, q1, v0, v1, tx, ty, px=None, py=None):
        self.q0, self.q1, self.v0, self.v1, self.tx, self.ty = (np.array(a, dtype=np.float64) for a in (q0, q1, v0, v1, tx, ty))
        if px is None:
            px, py = fk(self.q0, self.q1)
        self.px, self.py = np.array(px, dtype=np.float64), np.array(py, dtype=np.float64)
        self.step_count[:] = 0

    def obs(self):
        return make_obs(self.q0, self.q1, self.v0, self.v1, self.tx, self.ty, self.px, self.py)

    def step(self, a, auto_reset=True):
        """Returns (obs_after [N,11], reward [N], done [N] bool).  With auto_reset, finished envs return the reset obs
        (the reference caller does `if new: ob = env.reset()`, mlp_train.py:137-139)."""
        a = np.asarray(a, dtype=np.float64).reshape(self.n, 2)
        dx, dy = self.px - self.tx, self.py - self.ty
        rew = -np.sqrt(dx * dx + dy * dy) - (a[:, 0] ** 2 + a[:, 1] ** 2)   # stale fingertip, unclipped action
        u0, u1 = np.clip(a[:, 0], -1, 1), np.clip(a[:, 1], -1, 1)
        q0, q1, v0, v1 = self.q0, self.q1, self.v0, self.v1
        for _ in range(FRAME_SKIP):
            q0, q1, v0, v1, sq0, sq1 = rk4_substep(q0, q1, v0, v1, u0, u1)
        self.q0, self.q1, self.v0, self.v1 = q0, q1, v0, v1
        self.px, self.py = fk(sq0, sq1)
        self.step_count = self.step_count + 1
        done = self.step_count >= EPISODE_STEPS
        if auto_reset and done.any():
            self.episode = (self.episode + done.astype(np.uint32)).astype(np.uint32)
            self._reset_where(done)
        return self.obs(), rew, done
