"""Dev tool (CPU only): numpy float32 emulation of csrc/physics.cuh, to study the fp32 error of candidate arithmetic
variants against the float64 oracle on BASELINE config 2 (4096 envs x 500 steps, Philox random actions) before spending
GPU time.  Every device rounding is mirrored: fma(a,b,c) = float32(float64(a)*float64(b) + float64(c)).

    python scripts/physics_f32_study.py [--envs 4096] [--steps 500] [--variants base,comp,...]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reacher_c as RC  # noqa: E402
from oracle import reacher_np as RN  # noqa: E402

f32, f64 = np.float32, np.float64


def F(x):
    return np.asarray(x, dtype=f32)


def fma(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def mul(a, b):
    return (F(a) * F(b)).astype(f32)


def add(a, b):
    return (F(a) + F(b)).astype(f32)


def sub(a, b):
    return (F(a) - F(b)).astype(f32)


A_d, B_d, C_d = 6.8252147596789472e-04, 2.1886428820008897e-04, 1.729551475556301e-04
M11, M00c, B, B2, C = f32(1.0 + C_d), f32(1.0 + A_d), f32(B_d), f32(2.0 * B_d), f32(C_d)
GEAR, LIMIT = f32(200.0), f32(3.0)
K_LIM, B_LIM, INVW0 = f32(2770.0831024930749), f32(105.26315789473685), f32(0.9998272280584255)
H, H2, H6 = f32(0.01), f32(0.005), f32(0.01 / 6.0)
L0, LT = f32(0.1), f32(0.11)


def sincos_cw(x):
    t = fma(x, f32(0.6366197466850281), f32(12582912.0))
    j = t.view(np.int32)
    q = sub(t, f32(12582912.0))
    r = fma(q, f32(-1.5707963705062866), x)
    r = fma(q, f32(4.371138828673793e-08), r)
    r = fma(q, f32(1.7151245100058819e-15), r)
    z = mul(r, r)
    p = fma(f32(-0.00019495566084515303), z, f32(0.008331977762281895))
    p = fma(p, z, f32(-0.16666650772094727))
    sn = fma(mul(p, z), r, r)
    g = fma(f32(2.4438377295155078e-05), z, f32(-0.0013887366512790322))
    g = fma(g, z, f32(0.04166664555668831))
    cs = fma(mul(g, z), z, fma(z, f32(-0.5), f32(1.0)))
    swap = (j & 1) != 0
    ss, cc = np.where(swap, cs, sn), np.where(swap, sn, cs)
    s = np.where((j & 2) != 0, -ss, ss).astype(f32)
    c = np.where(((j + 1) & 2) != 0, -cc, cc).astype(f32)
    return s, c


# rotation of a cached (sin, cos) pair by a small increment d (|d| < ~0.5): Taylor polynomials, no range reduction
S3, S5, S7, S9 = f32(-1.0 / 6), f32(1.0 / 120), f32(-1.0 / 5040), f32(1.0 / 362880)
K2, K4, K6, K8 = f32(-0.5), f32(1.0 / 24), f32(-1.0 / 720), f32(1.0 / 40320)


def rotate(s, c, d, deg=9):
    z = mul(d, d)
    if deg >= 9:
        p = fma(S9, z, S7)
        p = fma(p, z, S5)
    else:
        p = fma(S7, z, S5)
    p = fma(p, z, S3)
    sd = fma(mul(p, z), d, d)                        # sin d
    g = fma(K8, z, K6) if deg >= 9 else K6
    g = fma(g, z, K4)
    g = fma(g, z, K2)
    cm = mul(g, z)                                   # cos d - 1
    s2 = fma(c, sd, fma(s, cm, s))
    c2 = fma(-s, sd, fma(c, cm, c))
    return s2, c2


class Opts:
    comp = False        # two-float accumulation of q0, q1
    rot = False         # stage trig by rotation of the cached pair
    poly_det = False    # 1/det as a polynomial in c1 instead of rcp.approx
    exact_div = False   # (study) exact division instead of rcp
    ptrick = False      # limit penetration of the RK4 stages from the exact substep-start value + the small increment


# 1/det(c1): det = (1+a+2b c)(1+c_) - (c_ + b c)^2 = d0 + d1 c + d2 c^2
_d0 = (1 + A_d) * (1 + C_d) - C_d * C_d
_d1 = 2 * B_d * (1 + C_d) - 2 * C_d * B_d
_d2 = -B_d * B_d
# 1/det = 1/d0 * 1/(1 + e),  e = (d1 c + d2 c^2)/d0 ; series to e^3, collected in powers of c (|c| <= 1)
_e1, _e2 = _d1 / _d0, _d2 / _d0
ID0 = f32(1.0 / _d0)
ID1 = f32(-_e1 / _d0)
ID2 = f32((-_e2 + _e1 * _e1) / _d0)
ID3 = f32((2 * _e1 * _e2 - _e1 ** 3) / _d0)


def accel(q1, s1, c1, v0, v1, g0, g1, o, over=None):
    m00, m01 = fma(B2, c1, M00c), fma(B, c1, C)
    bs = mul(B, s1)
    t0 = fma(mul(bs, v1), fma(f32(2.0), v0, v1), sub(g0, v0))
    t1 = fma(-mul(bs, v0), v0, sub(g1, v1))
    if o.poly_det:
        idet = fma(fma(fma(ID3, c1, ID2), c1, ID1), c1, ID0)
    elif o.exact_div:
        idet = (f64(1.0) / fma(m00, M11, -mul(m01, m01)).astype(f64)).astype(f32)
    else:
        idet = (f32(1.0) / fma(m00, M11, -mul(m01, m01))).astype(f32)
    a0 = mul(fma(M11, t0, -mul(m01, t1)), idet)
    a1 = mul(fma(m00, t1, -mul(m01, t0)), idet)
    if over is None:
        over = sub(np.abs(q1), LIMIT)
    act = over > 0
    if act.any():
        sgn = np.where(q1 > 0, f32(-1.0), f32(1.0)).astype(f32)
        x = np.minimum(mul(over, f32(1000.0)), f32(1.0))
        omx = sub(f32(1.0), x)
        y = np.where(x <= f32(0.5), mul(f32(2.0), mul(x, x)), fma(f32(-2.0), mul(omx, omx), f32(1.0))).astype(f32)
        imp = fma(f32(0.05), y, f32(0.9))
        aref = fma(K_LIM, mul(imp, over), -mul(B_LIM, mul(sgn, v1)))
        mi01, mi11 = -mul(m01, idet), mul(m00, idet)
        R = mul((sub(f32(1.0), imp) / imp).astype(f32), INVW0)
        f = (sub(aref, mul(sgn, a1)) / add(mi11, R)).astype(f32)
        f = np.maximum(f, f32(0.0))
        sf = mul(sgn, f)
        a0 = np.where(act, fma(mi01, sf, a0), a0).astype(f32)
        a1 = np.where(act, fma(mi11, sf, a1), a1).astype(f32)
    return a0, a1


def substep(e, g0, g1, s1, c1, o):
    q0, q1, v0, v1 = e["q0"], e["q1"], e["v0"], e["v1"]
    p0 = pa = pb = pc = None
    if o.ptrick:
        sg = np.where(q1 < 0, f32(-1.0), f32(1.0)).astype(f32)
        p0 = sub(np.abs(q1), LIMIT)                      # exact (Sterbenz) near the limit
        if o.comp:
            p0 = fma(sg, e["q1l"], p0)
        pa, pb = fma(mul(sg, H2), v1, p0), None
    f00, f01 = accel(q1, s1, c1, v0, v1, g0, g1, o, p0)
    qa1, va0, va1 = fma(H2, v1, q1), fma(H2, f00, v0), fma(H2, f01, v1)
    sa, ca = rotate(s1, c1, mul(H2, v1)) if o.rot else sincos_cw(qa1)
    f10, f11 = accel(qa1, sa, ca, va0, va1, g0, g1, o, pa)
    qb1, vb0, vb1 = fma(H2, va1, q1), fma(H2, f10, v0), fma(H2, f11, v1)
    if o.ptrick:
        pb = fma(mul(sg, H2), va1, p0)
    sb, cb = rotate(s1, c1, mul(H2, va1)) if o.rot else sincos_cw(qb1)
    f20, f21 = accel(qb1, sb, cb, vb0, vb1, g0, g1, o, pb)
    qc0, qc1, vc0, vc1 = fma(H, vb0, q0), fma(H, vb1, q1), fma(H, f20, v0), fma(H, f21, v1)
    if o.ptrick:
        pc = fma(mul(sg, H), vb1, p0)
    sc, cc = rotate(s1, c1, mul(H, vb1)) if o.rot else sincos_cw(qc1)
    f30, f31 = accel(qc1, sc, cc, vc0, vc1, g0, g1, o, pc)
    dq0 = mul(H6, add(add(v0, vc0), mul(f32(2.0), add(va0, vb0))))
    dq1 = mul(H6, add(add(v1, vc1), mul(f32(2.0), add(va1, vb1))))
    if o.comp:
        for nm, dq in (("q0", dq0), ("q1", dq1)):
            q, ql = e[nm], e[nm + "l"]
            d = add(dq, ql)                              # increment + carried low part
            s = add(q, d)
            ql2 = sub(d, sub(s, q))                      # Fast2Sum (|q| >= |d| except near zero, where the error is tiny anyway)
            e[nm], e[nm + "l"] = s, ql2
    else:
        e["q0"] = fma(H6, add(add(v0, vc0), mul(f32(2.0), add(va0, vb0))), q0)
        e["q1"] = fma(H6, add(add(v1, vc1), mul(f32(2.0), add(va1, vb1))), q1)
    e["v0"] = fma(H6, add(add(f00, f30), mul(f32(2.0), add(f10, f20))), v0)
    e["v1"] = fma(H6, add(add(f01, f31), mul(f32(2.0), add(f11, f21))), v1)
    return qc0, qc1, sc, cc


def fk_sc(s0, c0, s1, c1):
    c01, s01 = fma(c0, c1, -mul(s0, s1)), fma(s0, c1, mul(c0, s1))
    return fma(LT, c01, mul(L0, c0)), fma(LT, s01, mul(L0, s0))


def trig_of(e, nm, o):
    s, c = sincos_cw(e[nm])
    if o.comp:                                           # first-order correction by the low part: sin(q + l) = s + l c
        l = e[nm + "l"]
        s, c = fma(l, c, s), fma(-l, s, c)
    return s, c


def run(o, n, T, seed=0):
    ids = np.arange(n, dtype=np.uint32)
    orc = RC.ReacherOracleC(n, seed=seed)
    orc.reset()
    e = {}

    def reset_where(mask, episode):
        q0, q1, v0, v1, tx, ty = (F(a) for a in RN.reset_draws(seed, ids, episode))
        new = dict(q0=q0, q1=q1, v0=v0, v1=v1, tx=tx, ty=ty, q0l=np.zeros(n, f32), q1l=np.zeros(n, f32))
        s0, c0 = sincos_cw(q0); s1, c1 = sincos_cw(q1)
        px, py = fk_sc(s0, c0, s1, c1)
        new.update(px=px, py=py, s0=s0, c0=c0, s1=s1, c1=c1)
        for k, v in new.items():
            e[k] = v if k not in e else np.where(mask, v, e[k]).astype(f32)

    episode = np.zeros(n, np.uint32)
    step = np.zeros(n, np.int32)
    reset_where(np.ones(n, bool), episode)
    worst_obs = worst_rew = 0.0
    worst_hist = []
    for t in range(T):
        act = RN.random_actions(seed, ids, t)
        a0, a1 = act[:, 0], act[:, 1]
        dx, dy = sub(e["px"], e["tx"]), sub(e["py"], e["ty"])
        rew = -add(np.sqrt(fma(dx, dx, mul(dy, dy))).astype(f32), fma(a0, a0, mul(a1, a1)))
        g0, g1 = mul(GEAR, np.clip(a0, -1, 1)), mul(GEAR, np.clip(a1, -1, 1))
        s1, c1 = e["s1"], e["c1"]
        for sub_i in range(2):
            if sub_i:
                s1, c1 = trig_of(e, "q1", o)
            sq0, sq1, ss1, sc1 = substep(e, g0, g1, s1, c1, o)
        ss0, sc0 = sincos_cw(sq0)
        e["px"], e["py"] = fk_sc(ss0, sc0, ss1, sc1)
        e["s0"], e["c0"] = trig_of(e, "q0", o)
        e["s1"], e["c1"] = trig_of(e, "q1", o)
        step += 1
        done = step >= 50
        if done.any():
            episode = (episode + done.astype(np.uint32)).astype(np.uint32)
            reset_where(done, episode)
            step = np.where(done, 0, step)
        ob = np.stack([e["c0"], e["c1"], e["s0"], e["s1"], e["tx"], e["ty"], e["v0"], e["v1"], sub(e["px"], e["tx"]), sub(e["py"], e["ty"]),
                       np.zeros(n, f32)], -1).astype(f64)
        oref, rref, dref = orc.step(act.astype(f64))
        assert np.array_equal(done, dref)
        eo = np.abs(ob - oref) / np.maximum(1.0, np.abs(oref))
        er = np.abs(rew.astype(f64) - rref) / np.maximum(1.0, np.abs(rref))
        worst_hist.append(eo.max(axis=1))
        worst_obs, worst_rew = max(worst_obs, eo.max()), max(worst_rew, er.max())
    wh = np.stack(worst_hist)                       # [T, n]
    per_env = wh.max(axis=0)
    return worst_obs, worst_rew, per_env


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--variants", default="base,comp,rot,poly,comp+rot+poly,exact")
    a = ap.parse_args()
    for v in a.variants.split(","):
        o = Opts()
        o.comp, o.rot, o.poly_det, o.exact_div, o.ptrick = "comp" in v, "rot" in v, "poly" in v, "exact" in v, "ptrick" in v
        wo, wr, pe = run(o, a.envs, a.steps)
        q = np.quantile(pe, [0.5, 0.9, 0.99, 0.999])
        print("%-16s worst obs %.3g  worst rew %.3g   per-env worst: median %.2g  90%% %.2g  99%% %.2g  99.9%% %.2g  #>5e-5: %d"
              % (v, wo, wr, q[0], q[1], q[2], q[3], int((pe > 5e-5).sum())))


def locate(o, n, T, seed=0, thresh=1e-3):
    """debug: first (step, env, component) whose error exceeds thresh"""
    pass
