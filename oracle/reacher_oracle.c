/* ORACLE (test infrastructure, not product code) -- float64 C restatement of the Reacher-v2 hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Follows (reference = /root/reference, third-party arithmetic restated from its published model and PINNED by
 * the reference's recorded MuJoCo data src/distilation/tests/data/dataset.json, see oracle/reacher_np.py header):
 *   ro_reset_*   <- env.reset()  call sites src/distilation/mlp_train.py:112,138,200 (gym ReacherEnv.reset_model)
 *   ro_step      <- env.step(a)  call sites src/distilation/mlp_train.py:135,196 ; lstm_train.py:133,192
 *                   (gym ReacherEnv.step: reward from stale fingertip + unclipped action, do_simulation(a, 2),
 *                    MuJoCo-1.50 mj_step with RK4, soft joint limit, TimeLimit(50))
 *   ro_policy_fwd<- teacher.pi.pd.mean / pd.flat  src/distilation/teacher.py:12-16, run at mlp_train.py:123-125,165-167
 *                   (baselines MlpPolicy 2x64 tanh, obfilter clip +-5, state-independent logstd)
 *   ro_rollout_* <- teacher warm-up loop src/distilation/mlp_train.py:120-139 (vectorised over envs)
 * Twin of oracle/reacher_np.py; tests/test_oracle_c.py checks the two agree to 1e-13 and bit-exactly on RNG.
 *
 * Build: oracle/build.py (gcc -O2 -fopenmp -shared -fPIC) -> oracle/libreacher_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------- Philox4x32-10 ---------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void ro_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}
static inline float uniform_f32(uint32_t x, float lo, float hi) {
    float u = (float)(x >> 8) * 0x1p-24f;
    return fmaf(u, hi - lo, lo);
}

/* ---------------- model constants ---------------- */
#define PI 3.14159265358979323846
static const double RHO = 1000.0, RC = 0.01, LL = 0.1, L0 = 0.1, LT = 0.11;
static double A_, B_, C_, INVW0, K_LIM, B_LIM;
static int consts_ready = 0;
static void init_consts(void) {
    if (consts_ready) return;
    double m_link = RHO * PI * RC * RC * (LL + RC);
    double i_link = m_link * (3 * RC * RC + (LL + RC) * (LL + RC)) / 12.0;
    double m_tip = RHO * (4.0 / 3.0) * PI * RC * RC * RC;
    double i_tip = 0.4 * m_tip * RC * RC;
    C_ = i_link + m_link * 0.05 * 0.05 + i_tip + m_tip * LT * LT;
    B_ = L0 * (m_link * 0.05 + m_tip * LT);
    A_ = i_link + m_link * 0.05 * 0.05 + C_ + (m_link + m_tip) * L0 * L0;
    double m00 = 1.0 + A_ + 2 * B_, m01 = C_ + B_, m11 = 1.0 + C_;
    INVW0 = m00 / (m00 * m11 - m01 * m01);
    K_LIM = 1.0 / (0.95 * 0.95 * 0.02 * 0.02);
    B_LIM = 2.0 / (0.95 * 0.02);
    consts_ready = 1;
}
void ro_constants(double* out) { init_consts(); out[0] = A_; out[1] = B_; out[2] = C_; out[3] = INVW0; out[4] = K_LIM; out[5] = B_LIM; }

/* smallest | |q1| - 3 | any RK4 stage of the current env step has seen: the soft joint limit switches on discontinuously at dist = 0
 * (MuJoCo activates the constraint with its damping term in full), so a stage that lands within rounding distance of the threshold is
 * where an fp32 trajectory may legitimately take the other branch.  Parity tests use it to set such episodes aside (ro_step_graze). */
static _Thread_local double tl_graze = 1e300;
static inline void accel(double q1, double v0, double v1, double u0, double u1, double* a0o, double* a1o) {
    double c1 = cos(q1), s1 = sin(q1);
    { double gz = fabs(fabs(q1) - 3.0); if (gz < tl_graze) tl_graze = gz; }
    double m00 = 1.0 + A_ + 2 * B_ * c1, m01 = C_ + B_ * c1, m11 = 1.0 + C_;
    double cor0 = -B_ * s1 * (2 * v0 * v1 + v1 * v1), cor1 = B_ * s1 * v0 * v0;
    double t0 = 200.0 * u0 - v0 - cor0, t1 = 200.0 * u1 - v1 - cor1;
    double det = m00 * m11 - m01 * m01;
    double a0 = (m11 * t0 - m01 * t1) / det, a1 = (m00 * t1 - m01 * t0) / det;
    double mi01 = -m01 / det, mi11 = m00 / det;
    for (int side = 0; side < 2; ++side) {
        double sgn = side ? 1.0 : -1.0;
        double dist = side ? (q1 + 3.0) : (3.0 - q1);
        if (dist < 0) {
            double x = fmin(1.0, fabs(dist) / 0.001);
            double y = x <= 0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x);
            double imp = 0.9 + 0.05 * y;
            double aref = -B_LIM * (sgn * v1) - K_LIM * imp * dist;
            double R = (1 - imp) / imp * INVW0;
            double f = fmax(0.0, (aref - sgn * a1) / (mi11 + R));
            a0 += mi01 * sgn * f;
            a1 += mi11 * sgn * f;
        }
    }
    *a0o = a0; *a1o = a1;
}

typedef struct { double q0, q1, v0, v1, tx, ty, px, py; } Env;

static inline void fk(double q0, double q1, double* px, double* py) {
    *px = L0 * cos(q0) + LT * cos(q0 + q1);
    *py = L0 * sin(q0) + LT * sin(q0 + q1);
}
static inline void substep(Env* e, double u0, double u1, double* sq0, double* sq1) {
    const double h = 0.01;
    double q0 = e->q0, q1 = e->q1, v0 = e->v0, v1 = e->v1;
    double f00, f01, f10, f11, f20, f21, f30, f31;
    accel(q1, v0, v1, u0, u1, &f00, &f01);
    double qa0 = q0 + h / 2 * v0, qa1 = q1 + h / 2 * v1, va0 = v0 + h / 2 * f00, va1 = v1 + h / 2 * f01;
    accel(qa1, va0, va1, u0, u1, &f10, &f11);
    double qb0 = q0 + h / 2 * va0, qb1 = q1 + h / 2 * va1, vb0 = v0 + h / 2 * f10, vb1 = v1 + h / 2 * f11;
    accel(qb1, vb0, vb1, u0, u1, &f20, &f21);
    double qc0 = q0 + h * vb0, qc1 = q1 + h * vb1, vc0 = v0 + h * f20, vc1 = v1 + h * f21;
    accel(qc1, vc0, vc1, u0, u1, &f30, &f31);
    e->q0 = q0 + h / 6 * (v0 + 2 * va0 + 2 * vb0 + vc0);
    e->q1 = q1 + h / 6 * (v1 + 2 * va1 + 2 * vb1 + vc1);
    e->v0 = v0 + h / 6 * (f00 + 2 * f10 + 2 * f20 + f30);
    e->v1 = v1 + h / 6 * (f01 + 2 * f11 + 2 * f21 + f31);
    *sq0 = qc0; *sq1 = qc1;
}
static inline void load_env(const double* st, int n, int i, Env* e) {
    e->q0 = st[0 * (size_t)n + i]; e->q1 = st[1 * (size_t)n + i]; e->v0 = st[2 * (size_t)n + i]; e->v1 = st[3 * (size_t)n + i];
    e->tx = st[4 * (size_t)n + i]; e->ty = st[5 * (size_t)n + i]; e->px = st[6 * (size_t)n + i]; e->py = st[7 * (size_t)n + i];
}
static inline void store_env(double* st, int n, int i, const Env* e) {
    st[0 * (size_t)n + i] = e->q0; st[1 * (size_t)n + i] = e->q1; st[2 * (size_t)n + i] = e->v0; st[3 * (size_t)n + i] = e->v1;
    st[4 * (size_t)n + i] = e->tx; st[5 * (size_t)n + i] = e->ty; st[6 * (size_t)n + i] = e->px; st[7 * (size_t)n + i] = e->py;
}
static inline void write_obs(const Env* e, double* ob) {
    ob[0] = cos(e->q0); ob[1] = cos(e->q1); ob[2] = sin(e->q0); ob[3] = sin(e->q1);
    ob[4] = e->tx; ob[5] = e->ty; ob[6] = e->v0; ob[7] = e->v1;
    ob[8] = e->px - e->tx; ob[9] = e->py - e->ty; ob[10] = 0.0;
}
static inline void reset_env(Env* e, uint64_t seed, uint32_t env_id, uint32_t episode) {
    uint32_t r0[4], r1[4];
    philox4x32_10(env_id, episode, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r0);
    philox4x32_10(env_id, episode, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r1);
    e->q0 = uniform_f32(r0[0], -0.1f, 0.1f); e->q1 = uniform_f32(r0[1], -0.1f, 0.1f);
    e->tx = uniform_f32(r0[2], -0.2f, 0.2f); e->ty = uniform_f32(r0[3], -0.2f, 0.2f);
    e->v0 = uniform_f32(r1[0], -0.005f, 0.005f); e->v1 = uniform_f32(r1[1], -0.005f, 0.005f);
    fk(e->q0, e->q1, &e->px, &e->py);
}
/* one env step; returns reward; sets *done */
static inline double step_env(Env* e, double a0, double a1, int32_t* step, uint32_t* episode, uint64_t seed,
                              uint32_t env_id, int auto_reset, uint8_t* done) {
    double dx = e->px - e->tx, dy = e->py - e->ty;
    double rew = -sqrt(dx * dx + dy * dy) - (a0 * a0 + a1 * a1);
    double u0 = fmin(1.0, fmax(-1.0, a0)), u1 = fmin(1.0, fmax(-1.0, a1));
    double sq0, sq1;
    substep(e, u0, u1, &sq0, &sq1);
    substep(e, u0, u1, &sq0, &sq1);
    fk(sq0, sq1, &e->px, &e->py);
    *step += 1;
    *done = (uint8_t)(*step >= 50);
    if (*done && auto_reset) { *episode += 1; *step = 0; reset_env(e, seed, env_id, *episode); }
    return rew;
}

static void set_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}
int ro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* st: double[8][n] SoA (q0,q1,v0,v1,tx,ty,px,py); step int32[n]; episode uint32[n] */
void ro_reset_all(int n, uint64_t seed, uint32_t env_offset, double* st, int32_t* step, uint32_t* episode, double* obs, int nthreads) {
    init_consts(); set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        Env e; episode[i] = 0; step[i] = 0;
        reset_env(&e, seed, env_offset + (uint32_t)i, 0);
        store_env(st, n, i, &e);
        if (obs) write_obs(&e, obs + (size_t)i * 11);
    }
}
/* graze (may be NULL): per env, min over the 8 RK4 stage evaluations of this step of | |q1| - 3 | (see tl_graze) */
void ro_step_graze(int n, uint64_t seed, uint32_t env_offset, double* st, int32_t* step, uint32_t* episode, const double* act,
                   double* obs, double* rew, uint8_t* done, int auto_reset, int nthreads, double* graze) {
    init_consts(); set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        Env e; load_env(st, n, i, &e);
        tl_graze = 1e300;
        rew[i] = step_env(&e, act[2 * (size_t)i], act[2 * (size_t)i + 1], &step[i], &episode[i], seed, env_offset + (uint32_t)i,
                          auto_reset, &done[i]);
        if (graze) graze[i] = tl_graze;
        store_env(st, n, i, &e);
        write_obs(&e, obs + (size_t)i * 11);
    }
}
void ro_step(int n, uint64_t seed, uint32_t env_offset, double* st, int32_t* step, uint32_t* episode, const double* act,
             double* obs, double* rew, uint8_t* done, int auto_reset, int nthreads) {
    ro_step_graze(n, seed, env_offset, st, step, episode, act, obs, rew, done, auto_reset, nthreads, 0);
}
/* T steps with Philox random actions a~U(-1,1)^2 keyed (env, step0+t).
 * If traj != NULL it receives [T][n][12] = obs(11) + reward per step.  Returns sum of rewards (keeps work live). */
double ro_rollout_random(int n, uint64_t seed, uint32_t env_offset, double* st, int32_t* step, uint32_t* episode, int T,
                         uint32_t step0, double* traj, int nthreads) {
    init_consts(); set_threads(nthreads);
    double total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int i = 0; i < n; ++i) {
        Env e; load_env(st, n, i, &e);
        int32_t sc = step[i]; uint32_t ep = episode[i];
        for (int t = 0; t < T; ++t) {
            uint32_t r[4]; uint8_t d;
            philox4x32_10(env_offset + (uint32_t)i, step0 + (uint32_t)t, 0u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            double a0 = uniform_f32(r[0], -1.f, 1.f), a1 = uniform_f32(r[1], -1.f, 1.f);
            double rw = step_env(&e, a0, a1, &sc, &ep, seed, env_offset + (uint32_t)i, 1, &d);
            total += rw;
            if (traj) { double* o = traj + ((size_t)t * n + i) * 12; write_obs(&e, o); o[11] = rw; }
        }
        store_env(st, n, i, &e); step[i] = sc; episode[i] = ep;
    }
    return total;
}

/* ---------------- policy (baselines MlpPolicy 11->64->64->nout, tanh) ----------------
 * params (float32, flat): ob_mean[11] ob_std[11] W1[11][64] b1[64] W2[64][64] b2[64] W3[64][nout] b3[nout] logstd[2]
 * pdflat[4] = nout==2 ? (mean, logstd) : raw 4 outputs.  Arithmetic in double from float32 parameters. */
#define POL_H 64
static inline void policy_fwd1(const double* ob, const float* p, int nout, double* pdflat) {
    const float *mu = p, *sd = p + 11, *W1 = p + 22, *b1 = W1 + 11 * POL_H, *W2 = b1 + POL_H, *b2 = W2 + POL_H * POL_H,
                *W3 = b2 + POL_H, *b3 = W3 + POL_H * nout, *logstd = b3 + nout;
    double z[11], h1[POL_H], h2[POL_H];
    for (int k = 0; k < 11; ++k) { double v = (ob[k] - (double)mu[k]) / (double)sd[k]; z[k] = fmin(5.0, fmax(-5.0, v)); }
    for (int j = 0; j < POL_H; ++j) { double s = b1[j]; for (int k = 0; k < 11; ++k) s += z[k] * (double)W1[k * POL_H + j]; h1[j] = tanh(s); }
    for (int j = 0; j < POL_H; ++j) { double s = b2[j]; for (int k = 0; k < POL_H; ++k) s += h1[k] * (double)W2[k * POL_H + j]; h2[j] = tanh(s); }
    for (int j = 0; j < nout; ++j) { double s = b3[j]; for (int k = 0; k < POL_H; ++k) s += h2[k] * (double)W3[k * nout + j]; pdflat[j] = s; }
    if (nout == 2) { pdflat[2] = logstd[0]; pdflat[3] = logstd[1]; }
}
void ro_policy_fwd(int n, const double* obs, const float* params, int nout, double* pdflat, int nthreads) {
    set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) policy_fwd1(obs + (size_t)i * 11, params, nout, pdflat + (size_t)i * 4);
}
/* Teacher-in-the-loop rollout (mlp_train.py:120-139 vectorised): per step record (ob, pdflat, reward-of-this-step, done),
 * act with the pd mean.  Buffers may be NULL (timing only): obs_buf[T][n][11] pd_buf[T][n][4] rew_buf[T][n] done_buf[T][n].
 * The action is rounded to float32 (the device path carries actions in fp32; TF returns float32 too). */
double ro_rollout_policy(int n, uint64_t seed, uint32_t env_offset, double* st, int32_t* step, uint32_t* episode, int T,
                         const float* params, int nout, double* obs_buf, double* pd_buf, double* rew_buf, uint8_t* done_buf,
                         int nthreads) {
    init_consts(); set_threads(nthreads);
    double total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int i = 0; i < n; ++i) {
        Env e; load_env(st, n, i, &e);
        int32_t sc = step[i]; uint32_t ep = episode[i];
        for (int t = 0; t < T; ++t) {
            double ob[11], pd[4]; uint8_t d;
            write_obs(&e, ob);
            policy_fwd1(ob, params, nout, pd);
            double rw = step_env(&e, (double)(float)pd[0], (double)(float)pd[1], &sc, &ep, seed, env_offset + (uint32_t)i, 1, &d);
            total += rw;
            size_t r = (size_t)t * n + i;
            if (obs_buf) memcpy(obs_buf + r * 11, ob, sizeof ob);
            if (pd_buf) memcpy(pd_buf + r * 4, pd, sizeof pd);
            if (rew_buf) rew_buf[r] = rw;
            if (done_buf) done_buf[r] = d;
        }
        store_env(st, n, i, &e); step[i] = sc; episode[i] = ep;
    }
    return total;
}
