// Reacher-v2 dynamics for the device kernels: gym-0.10.5 reacher.xml under MuJoCo-1.50 semantics
// (RK4, h = 0.01, frame_skip 2, soft joint-1 limit, stale fingertip kinematics), as pinned by the reference's recorded
// trajectories (oracle/reacher_np.py is the float64 statement of the same model).
// Replaces env.step / env.reset at /root/reference src/distilation/mlp_train.py:112,135,138,196,200.
// Templated on the real type: float is the product path (state in registers, FP32 pipe); double is the strict mode.
#pragma once
#include "philox.cuh"

namespace rb {

template <typename T> struct Model {
    // c, b, a of the 2-link inertia matrix (armature 1 added on the diagonal); see oracle/reacher_np.py for the derivation
    static constexpr double A_d = 6.8252147596789472e-04, B_d = 2.1886428820008897e-04, C_d = 1.729551475556301e-04;
    static constexpr T M11 = (T)(1.0 + C_d);
    static constexpr T M00c = (T)(1.0 + A_d);
    static constexpr T B = (T)B_d, B2 = (T)(2.0 * B_d), C = (T)C_d;
    static constexpr T GEAR = (T)200.0;
    static constexpr T LIMIT = (T)3.0;
    static constexpr T K_LIM = (T)2770.0831024930749, B_LIM = (T)105.26315789473685, INVW0 = (T)0.9998272280584255;
    static constexpr T H = (T)0.01, H2 = (T)0.005, H6 = (T)(0.01 / 6.0);
    static constexpr T L0 = (T)0.1, LT = (T)0.11;
};

__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_t(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_t(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }

// joint accelerations; u already clipped to ctrlrange
template <typename T>
__device__ __forceinline__ void accel(T q1, T v0, T v1, T u0, T u1, T& a0, T& a1) {
    using M = Model<T>;
    T s1, c1;
    sincos_t(q1, &s1, &c1);
    const T m00 = M::M00c + M::B2 * c1, m01 = M::C + M::B * c1, m11 = M::M11;
    const T bs = M::B * s1;
    const T t0 = M::GEAR * u0 - v0 + bs * (T(2) * v0 * v1 + v1 * v1);
    const T t1 = M::GEAR * u1 - v1 - bs * v0 * v0;
    const T idet = T(1) / (m00 * m11 - m01 * m01);
    a0 = (m11 * t0 - m01 * t1) * idet;
    a1 = (m00 * t1 - m01 * t0) * idet;
    const T over = (q1 > T(0) ? q1 : -q1) - M::LIMIT;   // > 0  <=>  limit violated, dist = -over
    if (over > T(0)) {
        const T sgn = q1 > T(0) ? T(-1) : T(1);         // constraint Jacobian J = [0, sgn]
        const T dist = -over;
        const T x = over * T(1000) < T(1) ? over * T(1000) : T(1);
        const T y = x <= T(0.5) ? T(2) * x * x : T(1) - T(2) * (T(1) - x) * (T(1) - x);
        const T imp = T(0.9) + T(0.05) * y;
        const T aref = -M::B_LIM * (sgn * v1) - M::K_LIM * imp * dist;
        const T mi01 = -m01 * idet, mi11 = m00 * idet;
        const T R = (T(1) - imp) / imp * M::INVW0;
        T f = (aref - sgn * a1) / (mi11 + R);
        f = f > T(0) ? f : T(0);
        a0 += mi01 * sgn * f;
        a1 += mi11 * sgn * f;
    }
}

template <typename T> struct EnvState {
    T q0, q1, v0, v1;   // joint angles / velocities
    T tx, ty;           // target
    T px, py;           // fingertip position as MuJoCo's xpos holds it (last RK4 stage of the previous step)
    int step;           // steps taken in the current episode, [0,50)
    uint32_t episode;   // index into this env's Philox reset stream
};

template <typename T> __device__ __forceinline__ void fk(T q0, T q1, T& px, T& py) {
    T s0, c0, s01, c01;
    sincos_t(q0, &s0, &c0);
    sincos_t(q0 + q1, &s01, &c01);
    px = Model<T>::L0 * c0 + Model<T>::LT * c01;
    py = Model<T>::L0 * s0 + Model<T>::LT * s01;
}

// one mj_step (RK4).  sq0/sq1 = qpos of the LAST stage (what xpos is computed from).
template <typename T> __device__ __forceinline__ void substep(EnvState<T>& e, T u0, T u1, T& sq0, T& sq1) {
    using M = Model<T>;
    const T q0 = e.q0, q1 = e.q1, v0 = e.v0, v1 = e.v1;
    T f00, f01, f10, f11, f20, f21, f30, f31;
    accel(q1, v0, v1, u0, u1, f00, f01);
    const T qa1 = q1 + M::H2 * v1, va0 = v0 + M::H2 * f00, va1 = v1 + M::H2 * f01;
    accel(qa1, va0, va1, u0, u1, f10, f11);
    const T qb1 = q1 + M::H2 * va1, vb0 = v0 + M::H2 * f10, vb1 = v1 + M::H2 * f11;
    accel(qb1, vb0, vb1, u0, u1, f20, f21);
    const T qc0 = q0 + M::H * vb0, qc1 = q1 + M::H * vb1, vc0 = v0 + M::H * f20, vc1 = v1 + M::H * f21;
    accel(qc1, vc0, vc1, u0, u1, f30, f31);
    e.q0 = q0 + M::H6 * (v0 + T(2) * va0 + T(2) * vb0 + vc0);
    e.q1 = q1 + M::H6 * (v1 + T(2) * va1 + T(2) * vb1 + vc1);
    e.v0 = v0 + M::H6 * (f00 + T(2) * f10 + T(2) * f20 + f30);
    e.v1 = v1 + M::H6 * (f01 + T(2) * f11 + T(2) * f21 + f31);
    sq0 = qc0; sq1 = qc1;
}

template <typename T> __device__ __forceinline__ void reset_env(EnvState<T>& e, uint32_t k0, uint32_t k1, uint32_t gid) {
    const uint4 r0 = philox4x32_10(gid, e.episode, 0u, STREAM_RESET, k0, k1);
    const uint4 r1 = philox4x32_10(gid, e.episode, 1u, STREAM_RESET, k0, k1);
    e.q0 = (T)uniform_f32(r0.x, -0.1f, 0.1f);
    e.q1 = (T)uniform_f32(r0.y, -0.1f, 0.1f);
    e.tx = (T)uniform_f32(r0.z, -0.2f, 0.2f);
    e.ty = (T)uniform_f32(r0.w, -0.2f, 0.2f);
    e.v0 = (T)uniform_f32(r1.x, -0.005f, 0.005f);
    e.v1 = (T)uniform_f32(r1.y, -0.005f, 0.005f);
    fk(e.q0, e.q1, e.px, e.py);
    e.step = 0;
}

// gym ReacherEnv.step + TimeLimit(50) + auto-reset.  Returns reward; done set when this step ended the episode.
template <typename T>
__device__ __forceinline__ T step_env(EnvState<T>& e, T a0, T a1, uint32_t k0, uint32_t k1, uint32_t gid, bool& done) {
    const T dx = e.px - e.tx, dy = e.py - e.ty;
    // stale fingertip, unclipped action; explicit fma/mul so every kernel that inlines this rounds identically
    const T rew = -sqrt_t(fma_t(dx, dx, mul_t(dy, dy))) - fma_t(a0, a0, mul_t(a1, a1));
    const T u0 = a0 < T(-1) ? T(-1) : (a0 > T(1) ? T(1) : a0);
    const T u1 = a1 < T(-1) ? T(-1) : (a1 > T(1) ? T(1) : a1);
    T sq0, sq1;
    substep(e, u0, u1, sq0, sq1);
    substep(e, u0, u1, sq0, sq1);
    fk(sq0, sq1, e.px, e.py);
    e.step += 1;
    done = e.step >= 50;
    if (done) { e.episode += 1u; reset_env(e, k0, k1, gid); }
    return rew;
}

// 11-d observation: [cos q0, cos q1, sin q0, sin q1, tx, ty, v0, v1, px-tx, py-ty, 0]
template <typename T> __device__ __forceinline__ void observe(const EnvState<T>& e, float* ob) {
    T s0, c0, s1, c1;
    sincos_t(e.q0, &s0, &c0);
    sincos_t(e.q1, &s1, &c1);
    ob[0] = (float)c0; ob[1] = (float)c1; ob[2] = (float)s0; ob[3] = (float)s1;
    ob[4] = (float)e.tx; ob[5] = (float)e.ty; ob[6] = (float)e.v0; ob[7] = (float)e.v1;
    ob[8] = (float)(e.px - e.tx); ob[9] = (float)(e.py - e.ty); ob[10] = 0.f;
}

}  // namespace rb
