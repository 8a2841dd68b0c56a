// Two-headed LSTM student of the reference's backup experiment (/root/reference src/distilation/backup/student_rollout.py:130-200 graph,
// :303-328 placeholders and total loss, :331-338 Adam): one shared LSTMCell over an unrolled window, and per unrolled step an UN-shared head
//     trunk  = tanh(dense_D(m_tau))                                       'lstm_dense'     (:158)     D = 128
//     reward = dense_1(tanh(dense_64(trunk)))                             'reward1_dense', 'reward_out' (:160-163); the graph in the checked-in
//                                                                         tfevents has three hidden reward layers 64-32-64 (reward_hid / _2hid / _3hid)
//     pdflat = dense_4(tanh(dense_64(trunk)))                             'action_dense', from_flat (:165-172)
//     loss   = KL(student || teacher) summed over [T,B,2] (:196-200)  +  sum (reward - reward_target)^2 (:328)
// Input row: [dropout(ob, keep_prob) (11) | stepped action (2)] (:286-288).  LSTMCell = tf.contrib.rnn.LSTMCell(NUM_UNITS): z = [x, m] W + b,
// gates i, j, f, o, forget_bias 1 (same cell as lstm.cu).  Two graph variants exist in the reference and both are built here:
//     carry = 0   the checked-in source: `output, next_state = cell(x_i, state)` never reassigns `state` (:156), so EVERY unrolled step starts
//                 from the fed initial state and final_state is the initial state (:191)
//     carry = 1   the graph recorded in src/~/reacher/data/viz/1/events.out.tfevents.* (tests/golden/graph_facts.json): the state is carried
// All widths are run-time (spec[]): the source ships NUM_UNITS = 1, STEPS_UNROLLED = 2, LSTM_BATCH_SIZE = 2 with 100 / 20 / 100 commented
// beside them (:38-50).  Every product is a call of the tcgen05 bf16x3 GEMM (gemm_tc.cu), the T un-shared heads as ONE batched GEMM per
// layer; bias gradients and loss sums are fixed-order reductions, so results are bit-reproducible.
// Flat parameters: W_l[13+U][4U] b_l[4U], then per step tau (creation order of the source): Wd[U][D] bd | Wr_1[D][r_1] br_1 ... Wr_n | Wro[r_n][1] bro |
// Wa[D][A] ba | Wp[A][4] bp.
#include "common.cuh"
#include "dagger_input.cuh"
#include "gemm_tc.cuh"
#include "loss.cuh"

namespace rb {

constexpr int L2_IN = 13;          // 11 observation + 2 action columns
constexpr int L2_MAX_R = 4;

struct L2Spec {
    int U, T, carry, D, A, nR, rh[L2_MAX_R];
    int XH, LDX, G;
    int64_t o_Wl, o_bl, head0, head_sz;                                   // absolute offsets of the shared cell; per-step block
    int64_t h_Wd, h_bd, h_Wr[L2_MAX_R], h_br[L2_MAX_R], h_Wro, h_bro, h_Wa, h_ba, h_Wp, h_bp;      // offsets inside a step's block
    int64_t P;
    size_t splitk_floats;
};

static int l2_parse(const int* spec, L2Spec& s) {
    RB_REQUIRE(spec != nullptr, "spec is NULL");
    s.U = spec[0]; s.T = spec[1]; s.carry = spec[2]; s.D = spec[3]; s.A = spec[4]; s.nR = spec[5];
    RB_REQUIRE(s.U >= 1 && s.U <= 256, "units out of range (1..256)");
    RB_REQUIRE(s.T >= 1 && s.T <= 64, "unrolled steps out of range (1..64)");
    RB_REQUIRE(s.carry == 0 || s.carry == 1, "carry_state must be 0 or 1");
    RB_REQUIRE(s.D >= 1 && s.D <= 1024 && s.A >= 1 && s.A <= 1024, "head width out of range");
    RB_REQUIRE(s.nR >= 1 && s.nR <= L2_MAX_R, "reward head: 1..4 hidden layers");
    for (int k = 0; k < L2_MAX_R; ++k) {
        s.rh[k] = k < s.nR ? spec[6 + k] : 0;
        RB_REQUIRE(k >= s.nR || (s.rh[k] >= 1 && s.rh[k] <= 1024), "reward hidden width out of range");
    }
    s.XH = L2_IN + s.U; s.LDX = (s.XH + 3) & ~3; s.G = 4 * s.U;
    s.o_Wl = 0; s.o_bl = (int64_t)s.XH * s.G; s.head0 = s.o_bl + s.G;
    int64_t o = 0;
    s.h_Wd = o; o += (int64_t)s.U * s.D; s.h_bd = o; o += s.D;
    int prev = s.D;
    for (int k = 0; k < s.nR; ++k) { s.h_Wr[k] = o; o += (int64_t)prev * s.rh[k]; s.h_br[k] = o; o += s.rh[k]; prev = s.rh[k]; }
    s.h_Wro = o; o += prev; s.h_bro = o; o += 1;
    s.h_Wa = o; o += (int64_t)s.D * s.A; s.h_ba = o; o += s.A;
    s.h_Wp = o; o += (int64_t)s.A * 4; s.h_bp = o; o += 4;
    s.head_sz = o;
    s.P = s.head0 + (int64_t)s.T * s.head_sz;
    // split-K partials (<= 32 slices; gemm_bf16x3 takes fewer slices when the scratch is smaller): the shared cell's weight gradient, and the
    // head weight gradients batched over the T steps (capped at 64 MB)
    size_t big = (size_t)s.U * s.D;
    big = max(big, (size_t)s.D * s.A); big = max(big, (size_t)s.D * s.rh[0]);
    for (int k = 1; k < s.nR; ++k) big = max(big, (size_t)s.rh[k - 1] * s.rh[k]);
    s.splitk_floats = max((size_t)32 * s.LDX * s.G, min((size_t)32 * big * s.T, (size_t)1 << 24));
    return RB_OK;
}

struct L2Ws {
    float *xh, *z, *dz, *dxh, *c, *hh, *dh, *dc, *trunk, *dtrunk, *rh[L2_MAX_R], *drh[L2_MAX_R], *drew, *a1, *da1, *dpd, *splitk, *loss_part, *colpart;
};
static size_t l2_ws_floats(const L2Spec& s, int64_t B) {
    const size_t R = (size_t)s.T * B;
    size_t per_row = 2 * (size_t)s.LDX + 2 * (size_t)s.G + 2 * (size_t)s.U + 2 * (size_t)s.D + 2 * (size_t)s.A + 4 + 1;
    for (int k = 0; k < s.nR; ++k) per_row += 2 * (size_t)s.rh[k];
    return per_row * R + (size_t)(s.T + 2) * B * s.U + s.splitk_floats + 2 * 1024 + COLPART_FLOATS + 64 * 4 + 256;
}
static void l2_ws_carve(const L2Spec& s, int64_t B, float* ws, L2Ws& w) {
    const size_t R = (size_t)s.T * B;
    float* p = ws;
    auto take = [&](size_t n) { float* q = p; p += (n + 3) & ~(size_t)3; return q; };
    w.xh = take(R * s.LDX); w.z = take(R * s.G); w.dz = take(R * s.G); w.dxh = take(R * s.LDX);
    w.c = take((size_t)(s.T + 1) * B * s.U); w.hh = take(R * s.U); w.dh = take(R * s.U); w.dc = take((size_t)B * s.U);
    w.trunk = take(R * s.D); w.dtrunk = take(R * s.D);
    for (int k = 0; k < s.nR; ++k) { w.rh[k] = take(R * s.rh[k]); w.drh[k] = take(R * s.rh[k]); }
    w.drew = take(R); w.a1 = take(R * s.A); w.da1 = take(R * s.A); w.dpd = take(R * 4);
    w.splitk = take(s.splitk_floats); w.loss_part = take(2 * 1024); w.colpart = take(COLPART_FLOATS);
}

__device__ __forceinline__ float l2_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// rows of [dropout(ob) | action | m_prev | 0-pad]; the initial (c, m): c -> c[0], m -> the m_prev columns of step 0 (carry) or of EVERY step
__global__ void k_lstm2_inputs(int64_t R, int64_t B, int U, int LDX, int carry, const float* __restrict__ ob, const float* __restrict__ action,
                               float keep_prob, uint32_t k0, uint32_t k1, uint32_t sample_id0, uint32_t iteration, const uint32_t* __restrict__ clock,
                               const float* __restrict__ init_state, float* __restrict__ xh, float* __restrict__ c0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (clock) iteration = clock[0];                 // device-side step clock (CUDA-graph replay, rb_lstm2_step)
    if (i < R) {
        float o[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) o[k] = __ldg(ob + i * 11 + k);
        float4 r4[4];
        mlp_input_row(o, keep_prob, k0, k1, sample_id0 + (uint32_t)i, iteration, make_float4(0.f, 0.f, 0.f, 0.f), 0.f, r4);
        const float* r = reinterpret_cast<const float*>(r4);
        float* x = xh + i * LDX;
#pragma unroll
        for (int k = 0; k < 11; ++k) x[k] = r[k];
        x[11] = __ldg(action + i * 2); x[12] = __ldg(action + i * 2 + 1);
        for (int k = L2_IN + U; k < LDX; ++k) x[k] = 0.f;
    }
    const int64_t mrows = carry ? B : R;
    if (i < mrows * U) {
        const int64_t row = i / U, u = i - row * U, b = row % B;
        xh[row * LDX + L2_IN + u] = init_state ? __ldg(init_state + (B + b) * U + u) : 0.f;
        if (row < B) c0[b * U + u] = init_state ? __ldg(init_state + b * U + u) : 0.f;
    }
}

// gates of `rows` rows activated in place; c_prev of row r is c_prev[(r % B)]; m -> hh and (optionally) the m_prev columns of the next step's rows
__global__ void k_lstm2_cell_fwd(int64_t rows, int64_t B, int U, int LDX, float* __restrict__ z, const float* __restrict__ c_prev, float* __restrict__ c_out,
                                 float* __restrict__ hh, float* __restrict__ xh_next) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * U) return;
    const int64_t r = idx / U, u = idx - r * U, b = r % B;
    float* zr = z + r * 4 * U;
    const float i = l2_sigmoid(zr[u]), j = tanhf(zr[U + u]), f = l2_sigmoid(zr[2 * U + u] + 1.0f), o = l2_sigmoid(zr[3 * U + u]);
    const float c = fmaf(f, c_prev[b * U + u], i * j), m = o * tanhf(c);
    zr[u] = i; zr[U + u] = j; zr[2 * U + u] = f; zr[3 * U + u] = o;
    c_out[idx] = c;
    hh[idx] = m;
    if (xh_next) xh_next[r * LDX + L2_IN + u] = m;
}

// one cell of the BPTT: dm = head part (+ recurrent part), running dc (NULL: no carried state, every step stands alone)
__global__ void k_lstm2_cell_bwd(int64_t rows, int64_t B, int U, int LDX, const float* __restrict__ gates, const float* __restrict__ c_prev,
                                 const float* __restrict__ c, const float* __restrict__ dh_head, const float* __restrict__ dxh_next, float* __restrict__ dc,
                                 float* __restrict__ dz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * U) return;
    const int64_t r = idx / U, u = idx - r * U, b = r % B;
    const float* g = gates + r * 4 * U;
    const float i = g[u], j = g[U + u], f = g[2 * U + u], o = g[3 * U + u];
    const float tc = tanhf(c[idx]);
    const float dm = dh_head[idx] + (dxh_next ? dxh_next[r * LDX + L2_IN + u] : 0.f);
    const float dct = (dc ? dc[idx] : 0.f) + dm * o * (1.f - tc * tc);
    float* d = dz + r * 4 * U;
    d[u] = dct * j * i * (1.f - i);
    d[U + u] = dct * i * (1.f - j * j);
    d[2 * U + u] = dct * c_prev[b * U + u] * f * (1.f - f);
    d[3 * U + u] = dm * tc * o * (1.f - o);
    if (dc) dc[idx] = dct * f;
}

// KL row loss + squared reward error, their gradients; fixed grid => fixed summation order.  partial[blk] = KL part, partial[1024 + blk] = reward part
__global__ void k_lstm2_loss(int64_t R, const float4* __restrict__ s, const float4* __restrict__ t, const float* __restrict__ rew, const float* __restrict__ rew_target,
                             int loss_kind, float4* __restrict__ ds, float* __restrict__ drew, float* __restrict__ partial) {
    __shared__ float red[2][8];
    float lk = 0.f, lr = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
        float4 d;
        lk += pd_loss_row(s[i], t[i], loss_kind, d);
        ds[i] = d;
        const float e = rew[i] - rew_target[i];
        lr += e * e;
        drew[i] = 2.f * e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lk += __shfl_xor_sync(0xffffffffu, lk, o); lr += __shfl_xor_sync(0xffffffffu, lr, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = lk; red[1][threadIdx.x >> 5] = lr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        partial[blockIdx.x] = a; partial[1024 + blockIdx.x] = b;
    }
}
__global__ void k_lstm2_loss_final(const float* __restrict__ partial, int n, float* __restrict__ out3) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < n; ++i) { a += partial[i]; b += partial[1024 + i]; }
        out3[0] = a + b; out3[1] = a; out3[2] = b;                       // total_loss (:328) | KL part | reward part
    }
}

struct L2Call {
    const float *params, *ob, *action, *t_pd, *rew_target, *init_state;
    float keep_prob; uint64_t seed; uint32_t sample_id0, iteration;
    int64_t B; int loss_kind, fwd_only;
    float *s_out, *rew_out, *final_state, *gradloss;
    const uint32_t* clock;
};

#define RB_TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

struct L2Bat { int n = 1; long long sA = 0, sB = 0, sC = 0, sBias = 0, sH = 0; };
static int l2_gemm(const float* A, int lda, int a_mn, const float* Bm, int ldb, int b_mn, float* C, int ldc, int M, int N, int K, const float* bias, int act,
                   int accumulate, const float* H, int ldh, const L2Spec& s, L2Ws& w, bool split, int sms, cudaStream_t st, const L2Bat& bt = L2Bat()) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = Bm; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.act = act; g.accumulate = accumulate; g.H = H; g.ldh = ldh;
    g.batch = bt.n; g.sA = bt.sA; g.sB = bt.sB; g.sC = bt.sC; g.sBias = bt.sBias; g.sH = bt.sH;
    return gemm_bf16x3(g, split ? w.splitk : nullptr, split ? s.splitk_floats : 0, sms, st);
}

static int l2_run(const L2Spec& s, const L2Call& c, float* ws, cudaStream_t st) {
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t B = c.B, R = (int64_t)s.T * B;
    const int Bi = (int)B, Ri = (int)R, U = s.U, T = s.T, G = s.G, D = s.D, A = s.A;
    L2Ws w;
    l2_ws_carve(s, B, ws, w);
    const float* P = c.params;
    const float* Wl = P + s.o_Wl; const float* bl = P + s.o_bl;
    const float* H0 = P + s.head0;                                                // step 0's head block; step tau at + tau * head_sz
    const long long HS = s.head_sz;
    // ---- inputs --------------------------------------------------------------------------------------------------------------------
    k_lstm2_inputs<<<(unsigned)((max(R, (s.carry ? B : R) * U) + 255) / 256), 256, 0, st>>>(R, B, U, s.LDX, s.carry, c.ob, c.action, c.keep_prob, (uint32_t)c.seed,
                                                                                            (uint32_t)(c.seed >> 32), c.sample_id0, c.iteration, c.clock, c.init_state, w.xh, w.c);
    RB_CUDA(cudaGetLastError());
    // ---- the cell over the window --------------------------------------------------------------------------------------------------
    if (s.carry) {
        for (int t = 0; t < T; ++t) {
            float* xh_t = w.xh + (size_t)t * B * s.LDX;
            float* z_t = w.z + (size_t)t * B * G;
            RB_TRY(l2_gemm(xh_t, s.LDX, 0, Wl, G, 1, z_t, G, Bi, G, s.XH, bl, 0, 0, nullptr, 0, s, w, false, sms, st));
            k_lstm2_cell_fwd<<<(unsigned)((B * U + 255) / 256), 256, 0, st>>>(B, B, U, s.LDX, z_t, w.c + (size_t)t * B * U, w.c + (size_t)(t + 1) * B * U,
                                                                               w.hh + (size_t)t * B * U, t + 1 < T ? xh_t + (size_t)B * s.LDX : nullptr);
            RB_CUDA(cudaGetLastError());
        }
    } else {        // every step reads the same initial state (student_rollout.py:156): one product and one cell pass over all T * B rows
        RB_TRY(l2_gemm(w.xh, s.LDX, 0, Wl, G, 1, w.z, G, Ri, G, s.XH, bl, 0, 0, nullptr, 0, s, w, false, sms, st));
        k_lstm2_cell_fwd<<<(unsigned)((R * U + 255) / 256), 256, 0, st>>>(R, B, U, s.LDX, w.z, w.c, w.c + (size_t)B * U, w.hh, nullptr);
        RB_CUDA(cudaGetLastError());
    }
    if (c.final_state) {         // tf.identity(state) (:191): the carried state, or -- in the source's graph -- the state that was fed
        const float* cf = s.carry ? w.c + (size_t)T * B * U : w.c;
        RB_CUDA(cudaMemcpyAsync(c.final_state, cf, sizeof(float) * B * U, cudaMemcpyDeviceToDevice, st));
        if (s.carry) RB_CUDA(cudaMemcpyAsync(c.final_state + B * U, w.hh + (size_t)(T - 1) * B * U, sizeof(float) * B * U, cudaMemcpyDeviceToDevice, st));
        else RB_CUDA(cudaMemcpy2DAsync(c.final_state + B * U, sizeof(float) * U, w.xh + L2_IN, sizeof(float) * s.LDX, sizeof(float) * U, (size_t)B, cudaMemcpyDeviceToDevice, st));
    }
    // ---- heads, batched over the T steps (un-shared weights) ----------------------------------------------------------------------------
    auto bat = [&](int ld_in, int ld_out) { L2Bat b; b.n = T; b.sA = B * ld_in; b.sB = HS; b.sC = B * ld_out; b.sBias = HS; return b; };
    RB_TRY(l2_gemm(w.hh, U, 0, H0 + s.h_Wd, D, 1, w.trunk, D, Bi, D, U, H0 + s.h_bd, 1, 0, nullptr, 0, s, w, false, sms, st, bat(U, D)));
    {
        const float* in = w.trunk; int din = D;
        for (int k = 0; k < s.nR; ++k) {
            RB_TRY(l2_gemm(in, din, 0, H0 + s.h_Wr[k], s.rh[k], 1, w.rh[k], s.rh[k], Bi, s.rh[k], din, H0 + s.h_br[k], 1, 0, nullptr, 0, s, w, false, sms, st,
                           bat(din, s.rh[k])));
            in = w.rh[k]; din = s.rh[k];
        }
        RB_TRY(l2_gemm(in, din, 0, H0 + s.h_Wro, 1, 1, c.rew_out, 1, Bi, 1, din, H0 + s.h_bro, 0, 0, nullptr, 0, s, w, false, sms, st, bat(din, 1)));
    }
    RB_TRY(l2_gemm(w.trunk, D, 0, H0 + s.h_Wa, A, 1, w.a1, A, Bi, A, D, H0 + s.h_ba, 1, 0, nullptr, 0, s, w, false, sms, st, bat(D, A)));
    RB_TRY(l2_gemm(w.a1, A, 0, H0 + s.h_Wp, 4, 1, c.s_out, 4, Bi, 4, A, H0 + s.h_bp, 0, 0, nullptr, 0, s, w, false, sms, st, bat(A, 4)));
    if (c.fwd_only) return RB_OK;
    // ---- loss --------------------------------------------------------------------------------------------------------------------------
    const unsigned lb = (unsigned)min((int64_t)1024, (R + 255) / 256);
    k_lstm2_loss<<<lb, 256, 0, st>>>(R, (const float4*)c.s_out, (const float4*)c.t_pd, c.rew_out, c.rew_target, c.loss_kind, (float4*)w.dpd, w.drew, w.loss_part);
    k_lstm2_loss_final<<<1, 32, 0, st>>>(w.loss_part, (int)lb, c.gradloss + s.P);
    RB_CUDA(cudaGetLastError());
    // ---- heads backward (batched) ---------------------------------------------------------------------------------------------------------
    float* Gr = c.gradloss;
    float* GH = Gr + s.head0;
    // [dW] = in^T dout (split-K, batched: step tau's block at + tau * head_sz); db = column sums of dout
    auto wgrad = [&](const float* in, int din, const float* dout, int dn, int64_t oW, int64_t ob_) -> int {
        L2Bat b; b.n = T; b.sA = B * din; b.sB = B * dn; b.sC = HS;
        RB_TRY(l2_gemm(in, din, 1, dout, dn, 1, GH + oW, dn, din, dn, Bi, nullptr, 0, 0, nullptr, 0, s, w, true, sms, st, b));
        return colsum(dout, dn, B, dn, T, B * (long long)dn, GH + ob_, HS, w.colpart, st);
    };
    // d(in) = dout W^T (* (1 - in^2) when the input is a tanh output), optionally added to what is already there
    auto dgrad = [&](const float* dout, int dn, int64_t oW, float* din_buf, int din, const float* tanh_in, int accumulate) -> int {
        L2Bat b; b.n = T; b.sA = B * dn; b.sB = HS; b.sC = B * din; b.sH = B * din;
        return l2_gemm(dout, dn, 0, H0 + oW, dn, 0, din_buf, din, Bi, din, dn, nullptr, 0, accumulate, tanh_in, din, s, w, false, sms, st, b);
    };
    // action head
    RB_TRY(wgrad(w.a1, A, w.dpd, 4, s.h_Wp, s.h_bp));
    RB_TRY(dgrad(w.dpd, 4, s.h_Wp, w.da1, A, w.a1, 0));
    RB_TRY(wgrad(w.trunk, D, w.da1, A, s.h_Wa, s.h_ba));
    RB_TRY(dgrad(w.da1, A, s.h_Wa, w.dtrunk, D, w.trunk, 0));
    // reward head
    {
        const float* dout = w.drew; int dn = 1; int64_t oW = s.h_Wro, ob_ = s.h_bro;
        for (int k = s.nR - 1; k >= 0; --k) {
            RB_TRY(wgrad(w.rh[k], s.rh[k], dout, dn, oW, ob_));
            RB_TRY(dgrad(dout, dn, oW, w.drh[k], s.rh[k], w.rh[k], 0));
            dout = w.drh[k]; dn = s.rh[k]; oW = s.h_Wr[k]; ob_ = s.h_br[k];
        }
        RB_TRY(wgrad(w.trunk, D, dout, dn, oW, ob_));
        RB_TRY(dgrad(dout, dn, oW, w.dtrunk, D, w.trunk, 1));                   // both heads meet at the trunk
    }
    RB_TRY(wgrad(w.hh, U, w.dtrunk, D, s.h_Wd, s.h_bd));
    RB_TRY(dgrad(w.dtrunk, D, s.h_Wd, w.dh, U, nullptr, 0));
    // ---- back-propagation through time --------------------------------------------------------------------------------------------------
    if (s.carry) {
        RB_CUDA(cudaMemsetAsync(w.dc, 0, sizeof(float) * B * U, st));
        for (int t = T - 1; t >= 0; --t) {
            float* dz_t = w.dz + (size_t)t * B * G;
            float* dxh_t = w.dxh + (size_t)t * B * s.LDX;
            k_lstm2_cell_bwd<<<(unsigned)((B * U + 255) / 256), 256, 0, st>>>(B, B, U, s.LDX, w.z + (size_t)t * B * G, w.c + (size_t)t * B * U, w.c + (size_t)(t + 1) * B * U,
                                                                               w.dh + (size_t)t * B * U, t + 1 < T ? dxh_t + (size_t)B * s.LDX : nullptr, w.dc, dz_t);
            RB_CUDA(cudaGetLastError());
            if (t > 0)       // d[x | m_prev] = dz W_l^T; only the m_prev columns are used (the input has no trainable producer)
                RB_TRY(l2_gemm(dz_t, G, 0, Wl, G, 0, dxh_t, s.LDX, Bi, s.XH, G, nullptr, 0, 0, nullptr, 0, s, w, false, sms, st));
        }
    } else {
        k_lstm2_cell_bwd<<<(unsigned)((R * U + 255) / 256), 256, 0, st>>>(R, B, U, s.LDX, w.z, w.c, w.c + (size_t)B * U, w.dh, nullptr, nullptr, w.dz);
        RB_CUDA(cudaGetLastError());
    }
    // ---- the shared cell's gradients over all T * B rows ----------------------------------------------------------------------------------
    RB_TRY(l2_gemm(w.xh, s.LDX, 1, w.dz, G, 1, Gr + s.o_Wl, G, s.XH, G, Ri, nullptr, 0, 0, nullptr, 0, s, w, true, sms, st));
    RB_TRY(colsum(w.dz, G, R, G, 1, 0, Gr + s.o_bl, 0, w.colpart, st));
    return RB_OK;
}

// loss_grad with the dropout iteration taken from a device-side clock: the body of rb_lstm2_step (lstm.cu owns the clock / Adam kernels and the graph)
int lstm2_loss_grad_clocked(const int* spec, const float* params, const float* ob, const float* action, const float* t_pd, const float* reward_target,
                            const float* init_state, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, int loss_kind, float* s_out,
                            float* reward_out, float* gradloss, void* workspace, const uint32_t* clock, cudaStream_t st) {
    L2Spec s;
    RB_TRY(l2_parse(spec, s));
    RB_REQUIRE(B > 0 && B * s.T < ((int64_t)1 << 24) && keep_prob > 0.f && pd_loss_kind_ok(loss_kind), "bad batch / keep_prob / loss kind");
    L2Call c{};
    c.params = params; c.ob = ob; c.action = action; c.t_pd = t_pd; c.rew_target = reward_target; c.init_state = init_state; c.keep_prob = keep_prob;
    c.seed = seed; c.sample_id0 = sample_id0; c.B = B; c.loss_kind = loss_kind; c.s_out = s_out; c.rew_out = reward_out; c.gradloss = gradloss; c.clock = clock;
    return l2_run(s, c, (float*)workspace, st);
}
int64_t lstm2_params(const int* spec) { L2Spec s; return l2_parse(spec, s) ? -1 : s.P; }

}  // namespace rb

using namespace rb;

extern "C" {

int64_t rb_lstm2_param_count(const int* spec) {
    L2Spec s;
    return l2_parse(spec, s) ? -1 : s.P;
}

int64_t rb_lstm2_workspace_bytes(const int* spec, int64_t batch) {
    L2Spec s;
    if (batch <= 0 || l2_parse(spec, s)) return -1;
    return (int64_t)(sizeof(float) * l2_ws_floats(s, batch));
}

/* sess.run((s_ac, final_state_combined))  backup/student_rollout.py:527-535 -- forward over the window, keep_prob = 1 */
int rb_lstm2_fwd(const int* spec, const float* params, const float* ob, const float* action, const float* init_state, int64_t B, float* s_out,
                 float* reward_out, float* final_state, void* workspace, void* stream) {
    RB_REQUIRE(params && ob && action && s_out && reward_out && workspace, "NULL argument");
    L2Spec s;
    RB_TRY(l2_parse(spec, s));
    RB_REQUIRE(B > 0 && B * s.T < ((int64_t)1 << 24), "bad batch");
    L2Call c{};
    c.params = params; c.ob = ob; c.action = action; c.init_state = init_state; c.keep_prob = 1.f; c.B = B; c.fwd_only = 1;
    c.s_out = s_out; c.rew_out = reward_out; c.final_state = final_state;
    return l2_run(s, c, (float*)workspace, (cudaStream_t)stream);
}

/* sess.run([loss, minimize_adam]) minus Adam  backup/student_rollout.py:508-522 -- forward, KL + squared reward error, BPTT.
 * gradloss[P + 3] = flat gradient | total loss | KL part | reward part */
int rb_lstm2_loss_grad(const int* spec, const float* params, const float* ob, const float* action, const float* t_pd, const float* reward_target,
                       const float* init_state, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, uint32_t iteration, int loss_kind,
                       float* s_out, float* reward_out, float* final_state, float* gradloss, void* workspace, void* stream) {
    RB_REQUIRE(params && ob && action && t_pd && reward_target && s_out && reward_out && gradloss && workspace, "NULL argument");
    L2Spec s;
    RB_TRY(l2_parse(spec, s));
    RB_REQUIRE(B > 0 && B * s.T < ((int64_t)1 << 24), "bad batch");
    RB_REQUIRE(keep_prob > 0.f, "keep_prob must be > 0");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    L2Call c{};
    c.params = params; c.ob = ob; c.action = action; c.t_pd = t_pd; c.rew_target = reward_target; c.init_state = init_state; c.keep_prob = keep_prob;
    c.seed = seed; c.sample_id0 = sample_id0; c.iteration = iteration; c.B = B; c.loss_kind = loss_kind; c.s_out = s_out; c.rew_out = reward_out;
    c.final_state = final_state; c.gradloss = gradloss;
    return l2_run(s, c, (float*)workspace, (cudaStream_t)stream);
}

}  // extern "C"
