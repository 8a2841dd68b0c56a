// Reacher-v2 dynamics for the device kernels: gym-0.10.5 reacher.xml under MuJoCo-1.50 semantics
// (RK4, h = 0.01, frame_skip 2, soft joint-1 limit, stale fingertip kinematics), as pinned by the reference's recorded
// trajectories (oracle/reacher_np.py is the float64 statement of the same model).
// Replaces env.step / env.reset at /root/reference src/distilation/mlp_train.py:112,135,138,196,200.
//
// fp32, state in registers.  Every rounding is spelled out with __fmaf_rn / __fmul_rn / __fadd_rn so that each kernel
// that inlines these functions (single step, fused rollouts, DAgger act) produces bit-identical trajectories whatever
// the surrounding code lets the compiler contract.  Per env-step: 10 sincos (Cody-Waite + minimax polynomials, FMA pipe
// only, no slow path / local memory), 8 MUFU.RCP, 1 MUFU.SQRT; the sin/cos of the current joint angles ride along in
// the state ("trig cache") so the observation costs nothing and the first RK4 stage reuses them.
#pragma once
#include "philox.cuh"

namespace rb {

struct Model {
    // c, b, a of the 2-link inertia matrix (armature 1 added on the diagonal); see oracle/reacher_np.py for the derivation
    static constexpr double A_d = 6.8252147596789472e-04, B_d = 2.1886428820008897e-04, C_d = 1.729551475556301e-04;
    static constexpr float M11 = (float)(1.0 + C_d);
    static constexpr float M00c = (float)(1.0 + A_d);
    static constexpr float B = (float)B_d, B2 = (float)(2.0 * B_d), C = (float)C_d;
    static constexpr float GEAR = 200.0f;
    static constexpr float LIMIT = 3.0f;
    static constexpr float K_LIM = (float)2770.0831024930749, B_LIM = (float)105.26315789473685, INVW0 = (float)0.9998272280584255;
    static constexpr float H = 0.01f, H2 = 0.005f, H6 = (float)(0.01 / 6.0);
    static constexpr float L0 = 0.1f, LT = 0.11f;
};

#define RB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define RB_MUL(a, b) __fmul_rn((a), (b))
#define RB_ADD(a, b) __fadd_rn((a), (b))
#define RB_SUB(a, b) __fsub_rn((a), (b))

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// sin and cos for |x| < 2^21: x = q*pi/2 + r (three-constant Cody-Waite with FMAs, q from the 1.5*2^23 trick), minimax
// polynomials on [-pi/4, pi/4]; max abs error 9e-8 (measured against float64 over +-50 rad).  Joint angles stay below
// ~40 rad inside an episode (|qvel| <= 38 rad/s for 1 s), far inside the valid range.
__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
    const float t = RB_FMA(x, 0.6366197466850281f, 12582912.0f);
    const int j = __float_as_int(t);
    const float q = RB_SUB(t, 12582912.0f);
    float r = RB_FMA(q, -1.5707963705062866f, x);
    r = RB_FMA(q, 4.371138828673793e-08f, r);
    r = RB_FMA(q, 1.7151245100058819e-15f, r);
    const float z = RB_MUL(r, r);
    float p = RB_FMA(-0.00019495566084515303f, z, 0.008331977762281895f);
    p = RB_FMA(p, z, -0.16666650772094727f);
    const float sn = RB_FMA(RB_MUL(p, z), r, r);
    float g = RB_FMA(2.4438377295155078e-05f, z, -0.0013887366512790322f);
    g = RB_FMA(g, z, 0.04166664555668831f);
    const float cs = RB_FMA(RB_MUL(g, z), z, RB_FMA(z, -0.5f, 1.0f));
    const bool swap = (j & 1) != 0;
    const float ss = swap ? cs : sn, cc = swap ? sn : cs;
    s = __int_as_float(__float_as_int(ss) ^ ((j & 2) << 30));
    c = __int_as_float(__float_as_int(cc) ^ (((j + 1) & 2) << 30));
}

struct EnvState {
    float q0, q1, v0, v1;   // joint angles / velocities
    float tx, ty;           // target
    float px, py;           // fingertip position as MuJoCo's xpos holds it (last RK4 stage of the previous step)
    float s0, c0, s1, c1;   // trig cache: sin / cos of q0, q1 (always consistent with q0, q1)
    int step;               // steps taken in the current episode, [0,50)
    uint32_t episode;       // index into this env's Philox reset stream
};

__device__ __forceinline__ void refresh_trig(EnvState& e) {
    sincos_cw(e.q0, e.s0, e.c0);
    sincos_cw(e.q1, e.s1, e.c1);
}

// joint accelerations from sin/cos of q1; g = gear * clipped ctrl
__device__ __forceinline__ void accel(float q1, float s1, float c1, float v0, float v1, float g0, float g1, float& a0, float& a1) {
    using M = Model;
    const float m00 = RB_FMA(M::B2, c1, M::M00c), m01 = RB_FMA(M::B, c1, M::C);
    const float bs = RB_MUL(M::B, s1);
    const float t0 = RB_FMA(RB_MUL(bs, v1), RB_FMA(2.0f, v0, v1), RB_SUB(g0, v0));     // g0 - v0 + b s1 (2 v0 v1 + v1^2)
    const float t1 = RB_FMA(-RB_MUL(bs, v0), v0, RB_SUB(g1, v1));                      // g1 - v1 - b s1 v0^2
    const float idet = rcp_approx(RB_FMA(m00, M::M11, -RB_MUL(m01, m01)));
    a0 = RB_MUL(RB_FMA(M::M11, t0, -RB_MUL(m01, t1)), idet);
    a1 = RB_MUL(RB_FMA(m00, t1, -RB_MUL(m01, t0)), idet);
    const float over = RB_SUB(fabsf(q1), M::LIMIT);   // > 0  <=>  limit violated, dist = -over
    if (over > 0.0f) {
        const float sgn = q1 > 0.0f ? -1.0f : 1.0f;   // constraint Jacobian J = [0, sgn]
        const float x = fminf(RB_MUL(over, 1000.0f), 1.0f);
        const float omx = RB_SUB(1.0f, x);
        const float y = x <= 0.5f ? RB_MUL(2.0f, RB_MUL(x, x)) : RB_FMA(-2.0f, RB_MUL(omx, omx), 1.0f);
        const float imp = RB_FMA(0.05f, y, 0.9f);
        const float aref = RB_FMA(M::K_LIM, RB_MUL(imp, over), -RB_MUL(M::B_LIM, RB_MUL(sgn, v1)));   // -beta J.v - k imp dist
        const float mi01 = -RB_MUL(m01, idet), mi11 = RB_MUL(m00, idet);
        const float R = RB_MUL(__fdiv_rn(RB_SUB(1.0f, imp), imp), M::INVW0);
        float f = __fdiv_rn(RB_SUB(aref, RB_MUL(sgn, a1)), RB_ADD(mi11, R));
        f = fmaxf(f, 0.0f);
        const float sf = RB_MUL(sgn, f);
        a0 = RB_FMA(mi01, sf, a0);
        a1 = RB_FMA(mi11, sf, a1);
    }
}

// one mj_step (RK4).  In: (s1, c1) = sin/cos of e.q1.  Out: q/v advanced; (sq0, sq1) = qpos of the LAST stage (what xpos is
// computed from) and (s1, c1) = sin/cos of sq1.
__device__ __forceinline__ void substep(EnvState& e, float g0, float g1, float& s1, float& c1, float& sq0, float& sq1) {
    using M = Model;
    const float q0 = e.q0, q1 = e.q1, v0 = e.v0, v1 = e.v1;
    float f00, f01, f10, f11, f20, f21, f30, f31;
    accel(q1, s1, c1, v0, v1, g0, g1, f00, f01);
    const float qa1 = RB_FMA(M::H2, v1, q1), va0 = RB_FMA(M::H2, f00, v0), va1 = RB_FMA(M::H2, f01, v1);
    sincos_cw(qa1, s1, c1);
    accel(qa1, s1, c1, va0, va1, g0, g1, f10, f11);
    const float qb1 = RB_FMA(M::H2, va1, q1), vb0 = RB_FMA(M::H2, f10, v0), vb1 = RB_FMA(M::H2, f11, v1);
    sincos_cw(qb1, s1, c1);
    accel(qb1, s1, c1, vb0, vb1, g0, g1, f20, f21);
    const float qc0 = RB_FMA(M::H, vb0, q0), qc1 = RB_FMA(M::H, vb1, q1), vc0 = RB_FMA(M::H, f20, v0), vc1 = RB_FMA(M::H, f21, v1);
    sincos_cw(qc1, s1, c1);
    accel(qc1, s1, c1, vc0, vc1, g0, g1, f30, f31);
    e.q0 = RB_FMA(M::H6, RB_ADD(RB_ADD(v0, vc0), RB_MUL(2.0f, RB_ADD(va0, vb0))), q0);
    e.q1 = RB_FMA(M::H6, RB_ADD(RB_ADD(v1, vc1), RB_MUL(2.0f, RB_ADD(va1, vb1))), q1);
    e.v0 = RB_FMA(M::H6, RB_ADD(RB_ADD(f00, f30), RB_MUL(2.0f, RB_ADD(f10, f20))), v0);
    e.v1 = RB_FMA(M::H6, RB_ADD(RB_ADD(f01, f31), RB_MUL(2.0f, RB_ADD(f11, f21))), v1);
    sq0 = qc0; sq1 = qc1;
}

// fingertip from sin/cos of q0 and of q1 (angle addition for q0 + q1)
__device__ __forceinline__ void fk_sc(float s0, float c0, float s1, float c1, float& px, float& py) {
    const float c01 = RB_FMA(c0, c1, -RB_MUL(s0, s1)), s01 = RB_FMA(s0, c1, RB_MUL(c0, s1));
    px = RB_FMA(Model::LT, c01, RB_MUL(Model::L0, c0));
    py = RB_FMA(Model::LT, s01, RB_MUL(Model::L0, s0));
}
__device__ __forceinline__ void fk(float q0, float q1, float& px, float& py) {
    float s0, c0, s1, c1;
    sincos_cw(q0, s0, c0);
    sincos_cw(q1, s1, c1);
    fk_sc(s0, c0, s1, c1, px, py);
}

__device__ __forceinline__ void reset_env(EnvState& e, uint32_t k0, uint32_t k1, uint32_t gid) {
    const uint4 r0 = philox4x32_10(gid, e.episode, 0u, STREAM_RESET, k0, k1);
    const uint4 r1 = philox4x32_10(gid, e.episode, 1u, STREAM_RESET, k0, k1);
    e.q0 = uniform_f32(r0.x, -0.1f, 0.1f);
    e.q1 = uniform_f32(r0.y, -0.1f, 0.1f);
    e.tx = uniform_f32(r0.z, -0.2f, 0.2f);
    e.ty = uniform_f32(r0.w, -0.2f, 0.2f);
    e.v0 = uniform_f32(r1.x, -0.005f, 0.005f);
    e.v1 = uniform_f32(r1.y, -0.005f, 0.005f);
    refresh_trig(e);
    fk_sc(e.s0, e.c0, e.s1, e.c1, e.px, e.py);
    e.step = 0;
}

// gym ReacherEnv.step + TimeLimit(50) + auto-reset.  Returns reward; done set when this step ended the episode.
// Requires a valid trig cache on entry and leaves one on exit.
__device__ __forceinline__ float step_env(EnvState& e, float a0, float a1, uint32_t k0, uint32_t k1, uint32_t gid, bool& done) {
    const float dx = RB_SUB(e.px, e.tx), dy = RB_SUB(e.py, e.ty);
    // stale fingertip, unclipped action
    const float rew = -RB_ADD(sqrt_approx(RB_FMA(dx, dx, RB_MUL(dy, dy))), RB_FMA(a0, a0, RB_MUL(a1, a1)));
    const float g0 = RB_MUL(Model::GEAR, fminf(1.0f, fmaxf(-1.0f, a0)));
    const float g1 = RB_MUL(Model::GEAR, fminf(1.0f, fmaxf(-1.0f, a1)));
    float s1 = e.s1, c1 = e.c1, sq0, sq1;
#pragma unroll 1
    for (int sub = 0; sub < 2; ++sub) {              // frame_skip = 2; one copy of the RK4 body keeps the kernels inside the i-cache
        if (sub) sincos_cw(e.q1, s1, c1);
        substep(e, g0, g1, s1, c1, sq0, sq1);
    }
    float ss0, sc0;
    sincos_cw(sq0, ss0, sc0);
    fk_sc(ss0, sc0, s1, c1, e.px, e.py);      // xpos is left at the last RK4 stage of the last sub-step
    refresh_trig(e);
    e.step += 1;
    done = e.step >= 50;
    if (done) { e.episode += 1u; reset_env(e, k0, k1, gid); }
    return rew;
}

// 11-d observation: [cos q0, cos q1, sin q0, sin q1, tx, ty, v0, v1, px-tx, py-ty, 0]  (free: trig cache)
__device__ __forceinline__ void observe(const EnvState& e, float* ob) {
    ob[0] = e.c0; ob[1] = e.c1; ob[2] = e.s0; ob[3] = e.s1;
    ob[4] = e.tx; ob[5] = e.ty; ob[6] = e.v0; ob[7] = e.v1;
    ob[8] = RB_SUB(e.px, e.tx); ob[9] = RB_SUB(e.py, e.ty); ob[10] = 0.f;
}

// HBM state layout: qv float4[N] = (q0, q1, v0, v1), tp float4[N] = (tx, ty, px, py), ctr uint2[N] = (step, episode)
__device__ __forceinline__ EnvState load_state(const float4* qv, const float4* tp, const uint2* ctr, int64_t i) {
    const float4 a = qv[i], b = tp[i];
    const uint2 c = ctr[i];
    EnvState e;
    e.q0 = a.x; e.q1 = a.y; e.v0 = a.z; e.v1 = a.w; e.tx = b.x; e.ty = b.y; e.px = b.z; e.py = b.w;
    e.step = (int)c.x; e.episode = c.y;
    refresh_trig(e);
    return e;
}
__device__ __forceinline__ EnvState zero_state() {
    EnvState e;
    e.q0 = e.q1 = e.v0 = e.v1 = e.tx = e.ty = e.px = e.py = 0.f;
    e.s0 = e.s1 = 0.f; e.c0 = e.c1 = 1.f;
    e.step = 0; e.episode = 0;
    return e;
}
__device__ __forceinline__ void store_state(float4* qv, float4* tp, uint2* ctr, int64_t i, const EnvState& e) {
    qv[i] = make_float4(e.q0, e.q1, e.v0, e.v1);
    tp[i] = make_float4(e.tx, e.ty, e.px, e.py);
    ctr[i] = make_uint2((uint32_t)e.step, e.episode);
}

}  // namespace rb
