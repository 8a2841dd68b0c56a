// LSTM student (/root/reference src/distilation/student_nn.py:21-49, lstm_train.py:32-79): forward over a T = 10 step window, fused
// KL loss (loss.py:3-13) and full back-propagation through time, as a sequence of tensor-core GEMMs (gemm_tc.cu) and small
// element-wise kernels.  Graph, per window row (tau, b):
//     e      = prev_pdflat W_e + b_e                          tf.layers.dense(prev_pdflat, 32), shared by all steps   (:27)
//     x      = [dropout(ob, keep_prob) (11) | e (32)]                                                                 (:25,30)
//     z      = [x | m_{tau-1}] W_l + b_l ;  i, j, f, o = split(z, 4)      tf.contrib.rnn.LSTMCell(200), forget_bias 1   (:23,41)
//     c      = sigmoid(f + 1) c_{tau-1} + sigmoid(i) tanh(j) ;  m = sigmoid(o) tanh(c)
//     s_tau  = head_tau(m): 200 -> 64 -> 128 -> 64 -> 32 (tanh) -> 4, weights NOT shared between the unrolled steps      (:42-46)
// Flat parameter layout: W_e[4][32] b_e[32] W_l[243][800] b_l[800] then for tau = 0..9: W1[200][64] b1 W2[64][128] b2 W3[128][64] b3
// W4[64][32] b4 W5[32][4] b5.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dagger_input.cuh"
#include "gemm_tc.cuh"
#include "loss.cuh"
#include "lstm_recur.cuh"

namespace rb {

constexpr int LT = 10, LU = 200, LG = 800, LX = 43, LE = 32, LXH = 243, LDXH = 256;
// Activation buffers carry a constant ONES column right behind their last feature (hh: column 200 of 204, head layer outputs: column HD[l]
// of HD[l] + 4, xh: column 243 of 256).  The forward / dgrad GEMMs never read it (K = the feature count); the weight-gradient GEMM takes one
// more row of its transposed operand, and since every bias sits right behind its weight matrix in the parameter vector, that extra output
// row IS the bias gradient: no column-sum launches.
constexpr int LDHH = LU + 4;

constexpr int HD[6] = {200, 64, 128, 64, 32, 4};
#define HD_PAD(l) (HD[l] + 4)
static_assert(HD[1] == 64 && HD[2] == 128 && HD[3] == 64 && HD[4] == 32, "k_lstm_inputs spells the ones-column offsets out");
static inline int act_ld(int l) { return l == 0 ? LDHH : HD_PAD(l); }       // leading dimension of the input of head layer l
constexpr int L_WE = 0, L_BE = 128, L_WL = 160, L_BL = L_WL + LXH * LG, L_HEAD0 = L_BL + LG;
constexpr int L_HEAD_SZ = 200 * 64 + 64 + 64 * 128 + 128 + 128 * 64 + 64 + 64 * 32 + 32 + 32 * 4 + 4;      // 31652
constexpr int L_P = L_HEAD0 + LT * L_HEAD_SZ;                                                              // 511880
static inline int head_w_off(int l) { int o = 0; for (int i = 0; i < l; ++i) o += HD[i] * HD[i + 1] + HD[i + 1]; return o; }

// workspace layout (floats), R = T * B rows
struct LstmWs {
    float *xh, *z, *dz, *dxh, *c, *hh, *dh, *dc, *a[5], *da[5], *splitk, *colpart, *recur;
    size_t splitk_floats;
};
static size_t lstm_ws_floats(int64_t R, int64_t B, size_t* splitk) {
    size_t per_row = LDXH + LG + LG + LDXH + LDHH + LU;
    for (int l = 1; l <= 5; ++l) per_row += 2 * (size_t)HD[l] + 4;
    const size_t sk = (size_t)32 * LDXH * LG;            // split-K partials of the largest wgrad (243 x 800, <= 32 slices)
    if (splitk) *splitk = sk;
    return per_row * R + (size_t)(LT + 1) * B * LU + (size_t)B * LU + sk + 1024 + COLPART_FLOATS + lstm_recur_ws_floats(B) + 128;
}
static void lstm_ws_carve(float* ws, int64_t R, int64_t B, LstmWs& w) {
    float* p = ws;
    auto take = [&](size_t n) { float* q = p; p += (n + 3) & ~(size_t)3; return q; };
    w.xh = take(R * LDXH); w.z = take(R * LG); w.dz = take(R * LG); w.dxh = take(R * LDXH);
    w.c = take((size_t)(LT + 1) * B * LU); w.hh = take(R * LDHH); w.dh = take(R * LU); w.dc = take((size_t)B * LU);
    for (int l = 1; l <= 5; ++l) { w.a[l - 1] = take(R * HD_PAD(l)); w.da[l - 1] = take(R * HD[l]); }
    lstm_ws_floats(R, B, &w.splitk_floats);
    w.splitk = take(w.splitk_floats + 1024);
    w.colpart = take(COLPART_FLOATS);
    w.recur = take(lstm_recur_ws_floats(B));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// x rows: dropout(ob) -> xh[:, 0:11]; e = prev_pdflat W_e + b_e -> xh[:, 11:43]; initial m state -> xh rows of step 0, columns 43..242;
// initial c -> c[0].  The embedding is 4 fp32 FMAs per output (thread = one (row, column) pair: a warp reads one prev_pdflat row and one
// 128-byte row of W_e and writes 128 contiguous bytes) -- as a K = 4 tensor-core GEMM it was one more launch on the critical path.
__global__ void k_lstm_inputs(int64_t R, int64_t B, const float* __restrict__ ob, float keep_prob, uint32_t k0, uint32_t k1, uint32_t sample_id0,
                              uint32_t iteration, const uint32_t* __restrict__ clock, const float* __restrict__ init_state, float* __restrict__ xh,
                              float* __restrict__ c0, float* __restrict__ hh, float* __restrict__ a1, float* __restrict__ a2, float* __restrict__ a3,
                              float* __restrict__ a4, const float* __restrict__ prev_pd, const float* __restrict__ W_e, const float* __restrict__ b_e) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (clock) iteration = clock[0];                 // device-side step clock (CUDA-graph replay)
    if (i < R * LE) {
        const int64_t r = i / LE;
        const int n = (int)(i - r * LE);
        const float* p = prev_pd + r * 4;
        float e = __ldg(b_e + n);
#pragma unroll
        for (int k = 0; k < 4; ++k) e = fmaf(__ldg(p + k), __ldg(W_e + k * LE + n), e);
        xh[r * LDXH + 11 + n] = e;
    }
    if (i < R) {
        float o[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) o[k] = __ldg(ob + i * 11 + k);
        float4 r4[4];
        mlp_input_row(o, keep_prob, k0, k1, sample_id0 + (uint32_t)i, iteration, make_float4(0.f, 0.f, 0.f, 0.f), 0.f, r4);
        const float* r = reinterpret_cast<const float*>(r4);
#pragma unroll
        for (int k = 0; k < 11; ++k) xh[i * LDXH + k] = r[k];
#pragma unroll
        for (int k = LXH; k < LDXH; ++k) xh[i * LDXH + k] = k == LXH ? 1.f : 0.f;       // ones column -> b_l gradient
        hh[i * LDHH + LU] = 1.f;                                                         // ones columns of the head inputs -> bias gradients
        a1[i * 68 + 64] = 1.f; a2[i * 132 + 128] = 1.f; a3[i * 68 + 64] = 1.f; a4[i * 36 + 32] = 1.f;       // HD_PAD(l) / HD[l], l = 1..4
    }
    if (i < B * LU) {
        const int64_t b = i / LU, u = i - b * LU;
        c0[i] = init_state ? __ldg(init_state + i) : 0.f;
        xh[b * LDXH + LX + u] = init_state ? __ldg(init_state + B * LU + i) : 0.f;
    }
}

// gates (pre-activation, [B,800]) -> activated in place; c, m; m also goes to the next step's xh rows and to hh
__global__ void k_lstm_cell_fwd(int64_t B, float* __restrict__ z, const float* __restrict__ c_prev, float* __restrict__ c_out, float* __restrict__ hh,
                                float* __restrict__ xh_next) {       // hh rows have LDHH floats
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * LU) return;
    const int64_t b = idx / LU, u = idx - b * LU;
    float* zr = z + b * LG;
    const float i = sigmoidf_(zr[u]), j = tanhf(zr[LU + u]), f = sigmoidf_(zr[2 * LU + u] + 1.0f), o = sigmoidf_(zr[3 * LU + u]);
    const float c = fmaf(f, c_prev[idx], i * j), m = o * tanhf(c);
    zr[u] = i; zr[LU + u] = j; zr[2 * LU + u] = f; zr[3 * LU + u] = o;
    c_out[idx] = c;
    hh[b * LDHH + u] = m;
    if (xh_next) xh_next[b * LDXH + LX + u] = m;
}

// BPTT through one cell: dm (head part + recurrent part) and dc (running) -> dz (pre-activation gradients), dc for the previous step
__global__ void k_lstm_cell_bwd(int64_t B, const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c,
                                const float* __restrict__ dh_head, const float* __restrict__ dxh_next, float* __restrict__ dc, float* __restrict__ dz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * LU) return;
    const int64_t b = idx / LU, u = idx - b * LU;
    const float* g = gates + b * LG;
    const float i = g[u], j = g[LU + u], f = g[2 * LU + u], o = g[3 * LU + u];
    const float tc = tanhf(c[idx]);
    const float dm = dh_head[idx] + (dxh_next ? dxh_next[b * LDXH + LX + u] : 0.f);
    const float dct = dc[idx] + dm * o * (1.f - tc * tc);
    float* d = dz + b * LG;
    d[u] = dct * j * i * (1.f - i);
    d[LU + u] = dct * i * (1.f - j * j);
    d[2 * LU + u] = dct * c_prev[idx] * f * (1.f - f);
    d[3 * LU + u] = dm * tc * o * (1.f - o);
    dc[idx] = dct * f;
}

// KL(student || teacher) (or reverse) summed, and dL/ds, rows = T*B
__global__ void k_lstm_kl(int64_t R, const float4* __restrict__ s, const float4* __restrict__ t, int loss_kind, float4* __restrict__ ds,
                          float* __restrict__ loss_partial) {
    __shared__ float red[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < R) {
        float4 d;
        l = pd_loss_row(s[i], t[i], loss_kind, d);
        ds[i] = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        loss_partial[blockIdx.x] = tot;
    }
}
struct LstmCall {
    const float* params; const float* ob; const float* prev_pd; const float* t_pd; const float* init_state;
    float keep_prob; uint64_t seed; uint32_t sample_id0, iteration;
    int64_t B; int loss_kind, fwd_only;
    float* s_out; float* final_state; float* gradloss;
    const uint32_t* clock;
    int per_step_gemm;       // 1: the recurrence as one GEMM + one cell kernel per step (the first version; kept as a cross-check of lstm_recur.cu)
};

struct Bat { int n = 1; long long sA = 0, sB = 0, sC = 0, sBias = 0, sH = 0; };
static int gemm(const float* A, int lda, int a_mn, const float* Bm, int ldb, int b_mn, float* C, int ldc, int M, int N, int K, const float* bias, int act,
                int accumulate, const float* H, int ldh, LstmWs& w, int sms, cudaStream_t st, bool allow_split = false, const Bat& bt = Bat()) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = Bm; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.act = act; g.accumulate = accumulate; g.H = H; g.ldh = ldh;
    g.batch = bt.n; g.sA = bt.sA; g.sB = bt.sB; g.sC = bt.sC; g.sBias = bt.sBias; g.sH = bt.sH;
    return gemm_bf16x3(g, allow_split ? w.splitk : nullptr, allow_split ? w.splitk_floats : 0, sms, st);
}
#define RB_TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// Side stream for the head weight-gradient GEMMs: dW_l only needs dZ_l and the saved layer input, so the five [wgrad_l + split-K sum] pairs
// run beside the dgrad chain (and the start of the BPTT kernel) instead of between its links.  Fork / join with events, so the same
// code is captured as a two-branch CUDA graph by rb_lstm_step and runs as two real streams otherwise.  One set per device.
struct LstmSide { cudaStream_t s = nullptr; cudaEvent_t ready[5] = {}; cudaEvent_t done = nullptr; };
static int lstm_side(int device, LstmSide** out) {
    static LstmSide side[16];
    RB_REQUIRE(device >= 0 && device < 16, "device ordinal out of range");
    LstmSide& x = side[device];
    if (!x.s) {
        RB_CUDA(cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking));
        for (int i = 0; i < 5; ++i) RB_CUDA(cudaEventCreateWithFlags(&x.ready[i], cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming));
    }
    *out = &x;
    return RB_OK;
}

static int lstm_run(const LstmCall& c, float* ws, cudaStream_t st) {
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t B = c.B, R = LT * B;
    const int Bi = (int)B, Ri = (int)R;
    LstmWs w;
    lstm_ws_carve(ws, R, B, w);
    const float* P = c.params;
    constexpr int use_side = 1;           // weight-gradient GEMMs on a forked branch (two-branch CUDA graph): 0.578 -> 0.525 ms at 2048 windows
    LstmSide* side = nullptr;
    if (use_side) RB_TRY(lstm_side(device, &side));
    // ---- inputs: dropout(ob), initial state, embedding of prev_pdflat --------------------------------------------------------
    if (side) {                                                  // fork: the side branch starts from the same point as the main one
        RB_CUDA(cudaEventRecord(side->ready[0], st));
        RB_CUDA(cudaStreamWaitEvent(side->s, side->ready[0], 0));
    }
    static_assert(LE * LT >= LU, "the embedding's R * LE threads cover the B * LU state elements");
    k_lstm_inputs<<<(unsigned)((R * LE + 255) / 256), 256, 0, st>>>(R, B, c.ob, c.keep_prob, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), c.sample_id0,
                                                                   c.iteration, c.clock, c.init_state, w.xh, w.c, w.hh, w.a[0], w.a[1], w.a[2], w.a[3],
                                                                   c.prev_pd, P + L_WE, P + L_BE);
    RB_CUDA(cudaGetLastError());
    // ---- recurrence: one persistent cluster launch (lstm_recur.cu); RB_LSTM_RECUR=0 keeps the GEMM + cell launch sequence ------------
    const int use_recur = c.per_step_gemm ? 0 : 1;     // persistent recurrence kernels (default) or the GEMM + cell launch per step
    LstmRecurArgs ra{};
    ra.W_l = P + L_WL; ra.b_l = P + L_BL; ra.B = B; ra.xh = w.xh; ra.hh = w.hh; ra.hh_ld = LDHH; ra.c0 = w.c; ra.c_last = w.c + (size_t)LT * B * LU;
    ra.dh = w.dh; ra.dz = w.dz; ra.dxh = w.dxh; ra.scratch = w.recur;
    if (use_recur) {
        if (side) {                                              // W_l images only depend on the parameters: built beside the input kernels
            RB_TRY(lstm_recur_build_images(ra, side->s));
            RB_CUDA(cudaEventRecord(side->done, side->s));
            RB_CUDA(cudaStreamWaitEvent(st, side->done, 0));
        } else {
            RB_TRY(lstm_recur_build_images(ra, st));
        }
        RB_TRY(lstm_recur_forward(ra, st));
    } else {
        for (int t = 0; t < LT; ++t) {
            float* xh_t = w.xh + (size_t)t * B * LDXH;
            float* z_t = w.z + (size_t)t * B * LG;
            RB_TRY(gemm(xh_t, LDXH, 0, P + L_WL, LG, 1, z_t, LG, Bi, LG, LXH, P + L_BL, 0, 0, nullptr, 0, w, sms, st));
            k_lstm_cell_fwd<<<(unsigned)((B * LU + 255) / 256), 256, 0, st>>>(B, z_t, w.c + (size_t)t * B * LU, w.c + (size_t)(t + 1) * B * LU,
                                                                               w.hh + (size_t)t * B * LDHH, t + 1 < LT ? xh_t + (size_t)B * LDXH : nullptr);
            RB_CUDA(cudaGetLastError());
        }
    }
    if (side && !use_recur) {                                    // nothing ran on the side branch: still join it (stream capture needs every fork joined)
        RB_CUDA(cudaEventRecord(side->done, side->s));
        RB_CUDA(cudaStreamWaitEvent(st, side->done, 0));
    }
    if (c.final_state) {
        RB_CUDA(cudaMemcpyAsync(c.final_state, w.c + (size_t)LT * B * LU, sizeof(float) * B * LU, cudaMemcpyDeviceToDevice, st));
        RB_CUDA(cudaMemcpy2DAsync(c.final_state + B * LU, sizeof(float) * LU, w.hh + (size_t)(LT - 1) * B * LDHH, sizeof(float) * LDHH, sizeof(float) * LU,
                                  (size_t)B, cudaMemcpyDeviceToDevice, st));
    }
    // ---- per-step heads: the T heads have their own weights (student_nn.py:42-46) => one BATCHED GEMM per layer (batch = T) ------
    {
        const float* in = w.hh;
        for (int l = 0; l < 5; ++l) {
            const float* W = P + L_HEAD0 + head_w_off(l);
            float* out = l == 4 ? c.s_out : w.a[l];
            const int ld_in = act_ld(l), ld_out = l == 4 ? HD[5] : act_ld(l + 1);
            Bat bt; bt.n = LT; bt.sA = B * ld_in; bt.sB = L_HEAD_SZ; bt.sC = B * ld_out; bt.sBias = L_HEAD_SZ;
            RB_TRY(gemm(in, ld_in, 0, W, HD[l + 1], 1, out, ld_out, Bi, HD[l + 1], HD[l], W + HD[l] * HD[l + 1], l < 4 ? 1 : 0, 0, nullptr, 0, w, sms, st,
                        false, bt));
            in = out;
        }
    }
    if (c.fwd_only) return RB_OK;
    // ---- loss and dL/ds --------------------------------------------------------------------------------------------------------
    const unsigned kl_blocks = (unsigned)((R + 255) / 256);
    RB_REQUIRE(kl_blocks <= 1024, "window batch too large for the loss reduction scratch");
    float* loss_part = w.splitk + w.splitk_floats;       // 1024 spare floats behind the split-K area (lstm_ws_carve)
    k_lstm_kl<<<kl_blocks, 256, 0, st>>>(R, (const float4*)c.s_out, (const float4*)c.t_pd, c.loss_kind, (float4*)w.da[4], loss_part);
    RB_CUDA(cudaGetLastError());
    float* G = c.gradloss;
    // ---- heads backward, batched over the T steps ----------------------------------------------------------------------------------
    if (!side) RB_TRY(sum_serial(loss_part, (int)kl_blocks, c.gradloss + L_P, st));
    for (int l = 4; l >= 0; --l) {
        const float* W = P + L_HEAD0 + head_w_off(l);
        float* gW = G + L_HEAD0 + head_w_off(l);
        const float* dout = w.da[l];
        const float* in = l == 0 ? w.hh : w.a[l - 1];
        const int ld_in = act_ld(l);
        // [dW ; db] = [in | 1]^T dout : HD[l] + 1 output rows, the last one lands on the bias gradient (b follows W in the parameter vector)
        Bat bw; bw.n = LT; bw.sA = B * ld_in; bw.sB = B * HD[l + 1]; bw.sC = L_HEAD_SZ;
        cudaStream_t sw = st;
        if (side) {                                              // dZ_l is complete on the main stream here: the side stream may start dW_l
            RB_CUDA(cudaEventRecord(side->ready[l], st));
            RB_CUDA(cudaStreamWaitEvent(side->s, side->ready[l], 0));
            sw = side->s;
            if (l == 4) RB_TRY(sum_serial(loss_part, (int)kl_blocks, c.gradloss + L_P, sw));      // the serial loss sum leaves the critical path too
        }
        RB_TRY(gemm(in, ld_in, 1, dout, HD[l + 1], 1, gW, HD[l + 1], HD[l] + 1, HD[l + 1], Bi, nullptr, 0, 0, nullptr, 0, w, sms, sw, true, bw));
        float* din = l == 0 ? w.dh : w.da[l - 1];                                                     // d(in) = dout W^T (* tanh' of the layer input)
        Bat bd; bd.n = LT; bd.sA = B * HD[l + 1]; bd.sB = L_HEAD_SZ; bd.sC = B * HD[l]; bd.sH = B * ld_in;
        RB_TRY(gemm(dout, HD[l + 1], 0, W, HD[l + 1], 0, din, HD[l], Bi, HD[l], HD[l + 1], nullptr, 0, 0, l == 0 ? nullptr : in, ld_in, w, sms, st, false, bd));
    }
    if (side && !use_recur) {                                    // the GEMM-per-step fallback below uses the split-K scratch too: join first
        RB_CUDA(cudaEventRecord(side->done, side->s));
        RB_CUDA(cudaStreamWaitEvent(st, side->done, 0));
        side = nullptr;
    }
    // ---- back-propagation through time -------------------------------------------------------------------------------------------
    if (use_recur) {
        RB_TRY(lstm_recur_backward(ra, st));
    } else {
        RB_CUDA(cudaMemsetAsync(w.dc, 0, sizeof(float) * B * LU, st));
        for (int t = LT - 1; t >= 0; --t) {
            float* dz_t = w.dz + (size_t)t * B * LG;
            float* dxh_t = w.dxh + (size_t)t * B * LDXH;
            k_lstm_cell_bwd<<<(unsigned)((B * LU + 255) / 256), 256, 0, st>>>(B, w.z + (size_t)t * B * LG, w.c + (size_t)t * B * LU, w.c + (size_t)(t + 1) * B * LU,
                                                                               w.dh + (size_t)t * B * LU, t + 1 < LT ? dxh_t + (size_t)B * LDXH : nullptr, w.dc, dz_t);
            RB_CUDA(cudaGetLastError());
            RB_TRY(gemm(dz_t, LG, 0, P + L_WL, LG, 0, dxh_t, LDXH, Bi, LXH, LG, nullptr, 0, 0, nullptr, 0, w, sms, st, true));  // d[x | m_prev] = dz W_l^T
        }
    }
    if (side) {                                                  // join: the shared split-K scratch and the gradient vector are next used here
        RB_CUDA(cudaEventRecord(side->done, side->s));
        RB_CUDA(cudaStreamWaitEvent(st, side->done, 0));
    }
    // ---- weight gradients of the shared parts, over all T*B rows ---------------------------------------------------------------------
    // The W_l gradient keeps the first 24 slices' worth of the split-K scratch, the embedding gradient (+ its bias column sum) takes the rest and
    // runs on the side branch beside it.
    LstmWs wl = w, we = w;
    wl.splitk_floats = (size_t)24 * LDXH * LG;
    we.splitk = w.splitk + wl.splitk_floats; we.splitk_floats = w.splitk_floats - wl.splitk_floats;
    cudaStream_t se = st;
    if (side) {
        RB_CUDA(cudaEventRecord(side->ready[0], st));
        RB_CUDA(cudaStreamWaitEvent(side->s, side->ready[0], 0));
        se = side->s;
    }
    RB_TRY(gemm(w.xh, LDXH, 1, w.dz, LG, 1, G + L_WL, LG, LXH + 1, LG, Ri, nullptr, 0, 0, nullptr, 0, wl, sms, st, true));   // row 243 (ones column) = db_l
    RB_TRY(gemm(c.prev_pd, 4, 1, w.dxh + 11, LDXH, 1, G + L_WE, LE, 4, LE, Ri, nullptr, 0, 0, nullptr, 0, we, sms, se, true));
    RB_TRY(colsum(w.dxh + 11, LDXH, R, LE, 1, 0, G + L_BE, 0, w.colpart, se));
    RB_CUDA(cudaGetLastError());
    if (side) {
        RB_CUDA(cudaEventRecord(side->done, side->s));
        RB_CUDA(cudaStreamWaitEvent(st, side->done, 0));
    }
    return RB_OK;
}

// step clock helpers: lr_t of this step from clock[1] + 1 (same formula as the host, in double); advance after the update
__global__ void k_lstm_clock_prep(const uint32_t* clock, float lr, float b1, float b2, float* lr_t) {
    const double t = (double)(clock[1] + 1u);
    *lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
}
__global__ void k_lstm_clock_advance(uint32_t* clock) { clock[0] += 1u; clock[1] += 1u; }
__global__ void k_lstm_clock_set(uint32_t* clock, uint32_t iteration, uint32_t adam_t) { clock[0] = iteration; clock[1] = adam_t; }
__global__ void k_adam_dev_lr(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int64_t n,
                              const float* __restrict__ lr_t, float b1, float b2, float eps, float gscale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g[i], *lr_t, b1, b2, eps, gscale);
    p[i] = pi; m[i] = mi; v[i] = vi;
}

// implemented in lstm2.cu
int lstm2_loss_grad_clocked(const int* spec, const float* params, const float* ob, const float* action, const float* t_pd, const float* reward_target,
                            const float* init_state, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, int loss_kind, float* s_out,
                            float* reward_out, float* gradloss, void* workspace, const uint32_t* clock, cudaStream_t st);
int64_t lstm2_params(const int* spec);

}  // namespace rb

struct rb_lstm_ctx {
    int device = 0;
    uint32_t* clock = nullptr;       // {dropout iteration, adam step}
    float* lr_t = nullptr;
    cudaGraphExec_t gexec = nullptr;
    uint64_t gkey = 0;
    cudaStream_t cap_stream = nullptr;
};

using namespace rb;

extern "C" {

int64_t rb_lstm_param_count(void) { return L_P; }
int rb_lstm_steps(void) { return LT; }
int rb_lstm_units(void) { return LU; }
int64_t rb_lstm_workspace_bytes(int64_t batch) { return batch > 0 ? (int64_t)(sizeof(float) * lstm_ws_floats(LT * batch, batch, nullptr)) : -1; }

/* sess.run((s_pdflat_batch, final_state_batch))  lstm_train.py:171-182 -- forward over the T-step window from a given state */
int rb_lstm_fwd(const float* params, const float* ob, const float* prev_pd, const float* init_state, int64_t B, float* s_out, float* final_state,
                void* workspace, void* stream) {
    RB_REQUIRE(params && ob && prev_pd && s_out && workspace, "NULL argument");
    RB_REQUIRE(B > 0 && B * LT < ((int64_t)1 << 30), "bad batch");
    LstmCall c{};
    c.params = params; c.ob = ob; c.prev_pd = prev_pd; c.init_state = init_state; c.keep_prob = 1.f; c.B = B; c.fwd_only = 1;
    c.s_out = s_out; c.final_state = final_state;
    return lstm_run(c, (float*)workspace, (cudaStream_t)stream);
}

/* sess.run([loss, minimize_adam]) minus Adam  lstm_train.py:145-160 -- forward, KL loss, BPTT; gradloss[P+1] = flat gradient | loss */
int rb_lstm_loss_grad(const float* params, const float* ob, const float* prev_pd, const float* t_pd, const float* init_state, int64_t B, float keep_prob,
                      uint64_t seed, uint32_t sample_id0, uint32_t iteration, int loss_kind, float* s_out, float* final_state, float* gradloss,
                      void* workspace, void* stream) {
    RB_REQUIRE(params && ob && prev_pd && t_pd && s_out && gradloss && workspace, "NULL argument");
    RB_REQUIRE(B > 0 && B * LT < ((int64_t)1 << 30), "bad batch");
    RB_REQUIRE(keep_prob > 0.f, "keep_prob must be > 0");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    LstmCall c{};
    c.params = params; c.ob = ob; c.prev_pd = prev_pd; c.t_pd = t_pd; c.init_state = init_state; c.keep_prob = keep_prob; c.seed = seed;
    c.sample_id0 = sample_id0; c.iteration = iteration; c.B = B; c.loss_kind = loss_kind; c.s_out = s_out; c.final_state = final_state;
    c.gradloss = gradloss;
    return lstm_run(c, (float*)workspace, (cudaStream_t)stream);
}

/* One optimiser step of the LSTM student (loss_grad + TF-form Adam) with the per-step values (dropout iteration, Adam step) taken from a
 * device-side clock, so the ~250 launches are captured ONCE in a CUDA graph and replayed with one cudaGraphLaunch.               */
int rb_lstm_ctx_destroy(rb_lstm_ctx* c) {
    if (!c) return RB_OK;
    DeviceGuard guard(c->device);
    cudaFree(c->clock); cudaFree(c->lr_t);
    if (c->gexec) cudaGraphExecDestroy(c->gexec);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    delete c;
    return RB_OK;
}
int rb_lstm_ctx_create(rb_lstm_ctx** out, int device) {
    RB_REQUIRE(out != nullptr, "out is NULL");
    DeviceGuard guard(device);
    rb_lstm_ctx* c = new rb_lstm_ctx();
    c->device = device;
    cudaError_t err = cudaMalloc(&c->clock, 4 * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaMalloc(&c->lr_t, sizeof(float));
    if (err == cudaSuccess) err = cudaMemset(c->clock, 0, 4 * sizeof(uint32_t));
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) { rb_lstm_ctx_destroy(c); return cuda_fail(err, "rb_lstm_ctx_create"); }
    *out = c;
    return RB_OK;
}
int rb_lstm_ctx_set_clock(rb_lstm_ctx* c, uint32_t iteration, uint32_t adam_step, void* stream) {
    RB_REQUIRE(c != nullptr, "NULL argument");
    k_lstm_clock_set<<<1, 1, 0, (cudaStream_t)stream>>>(c->clock, iteration, adam_step);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}
int rb_lstm_step(rb_lstm_ctx* ctx, float* params, float* m, float* v, const float* ob, const float* prev_pd, const float* t_pd, const float* init_state,
                 int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, int loss_kind, float* s_out, float* gradloss, void* workspace,
                 float lr, float b1, float b2, float eps, float gscale, int use_graph, void* stream) {
    RB_REQUIRE(ctx && params && m && v && ob && prev_pd && t_pd && s_out && gradloss && workspace, "NULL argument");
    RB_REQUIRE(B > 0 && B * LT < ((int64_t)1 << 30) && keep_prob > 0.f, "bad batch / keep_prob");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    cudaStream_t st = (cudaStream_t)stream;
    auto issue = [&](cudaStream_t s) -> int {
        k_lstm_clock_prep<<<1, 1, 0, s>>>(ctx->clock, lr, b1, b2, ctx->lr_t);
        LstmCall c{};
        c.params = params; c.ob = ob; c.prev_pd = prev_pd; c.t_pd = t_pd; c.init_state = init_state; c.keep_prob = keep_prob; c.seed = seed;
        c.sample_id0 = sample_id0; c.B = B; c.loss_kind = loss_kind; c.s_out = s_out; c.gradloss = gradloss; c.clock = ctx->clock;
        int rc = lstm_run(c, (float*)workspace, s);
        if (rc) return rc;
        k_adam_dev_lr<<<(unsigned)((L_P + 255) / 256), 256, 0, s>>>(params, m, v, gradloss, L_P, ctx->lr_t, b1, b2, eps, gscale);
        k_lstm_clock_advance<<<1, 1, 0, s>>>(ctx->clock);
        RB_CUDA(cudaGetLastError());
        return RB_OK;
    };
    if (!use_graph) return issue(st);
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t vv) { key = (key ^ vv) * 1099511628211ull; };
    const void* ptrs[] = {params, m, v, ob, prev_pd, t_pd, init_state, s_out, gradloss, workspace};
    for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
    const float fl[] = {keep_prob, lr, b1, b2, eps, gscale};
    for (float f : fl) { uint32_t u; memcpy(&u, &f, 4); mix(u); }
    mix((uint64_t)B); mix(seed); mix(sample_id0); mix((uint64_t)loss_kind);
    if (!ctx->gexec || ctx->gkey != key) {
        if (ctx->gexec) { cudaGraphExecDestroy(ctx->gexec); ctx->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        RB_CUDA(cudaStreamSynchronize(st));
        RB_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeRelaxed));
        const int rc = issue(ctx->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        const cudaError_t ci = cudaGraphInstantiate(&ctx->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) { ctx->gexec = nullptr; return cuda_fail(ci, "cudaGraphInstantiate"); }
        ctx->gkey = key;
    }
    RB_CUDA(cudaGraphLaunch(ctx->gexec, st));
    return RB_OK;
}

/* rb_lstm2_loss_grad + TF-form Adam of the two-headed LSTM student as ONE CUDA-graph launch (same context type and device-side clock as
 * rb_lstm_step): at the reference's own sizes (100 windows x 20 steps, ~70 small launches) the step is launch-bound without it.      */
int rb_lstm2_step(rb_lstm_ctx* ctx, const int* spec, float* params, float* m, float* v, const float* ob, const float* action, const float* t_pd,
                  const float* reward_target, const float* init_state, int64_t B, float keep_prob, uint64_t seed, uint32_t sample_id0, int loss_kind,
                  float* s_out, float* reward_out, float* gradloss, void* workspace, float lr, float b1, float b2, float eps, float gscale, int use_graph,
                  void* stream) {
    RB_REQUIRE(ctx && spec && params && m && v && ob && action && t_pd && reward_target && s_out && reward_out && gradloss && workspace, "NULL argument");
    const int64_t P = lstm2_params(spec);
    RB_REQUIRE(P > 0, "bad spec");
    cudaStream_t st = (cudaStream_t)stream;
    auto issue = [&](cudaStream_t s) -> int {
        k_lstm_clock_prep<<<1, 1, 0, s>>>(ctx->clock, lr, b1, b2, ctx->lr_t);
        int rc = lstm2_loss_grad_clocked(spec, params, ob, action, t_pd, reward_target, init_state, B, keep_prob, seed, sample_id0, loss_kind, s_out, reward_out,
                                         gradloss, workspace, ctx->clock, s);
        if (rc) return rc;
        k_adam_dev_lr<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(params, m, v, gradloss, P, ctx->lr_t, b1, b2, eps, gscale);
        k_lstm_clock_advance<<<1, 1, 0, s>>>(ctx->clock);
        RB_CUDA(cudaGetLastError());
        return RB_OK;
    };
    if (!use_graph) return issue(st);
    uint64_t key = 1469598103934665603ull ^ 0x2ull;
    auto mix = [&](uint64_t vv) { key = (key ^ vv) * 1099511628211ull; };
    const void* ptrs[] = {params, m, v, ob, action, t_pd, reward_target, init_state, s_out, reward_out, gradloss, workspace};
    for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
    const float fl[] = {keep_prob, lr, b1, b2, eps, gscale};
    for (float f : fl) { uint32_t u; memcpy(&u, &f, 4); mix(u); }
    for (int i = 0; i < RB_LSTM2_SPEC_LEN; ++i) mix((uint64_t)(uint32_t)spec[i]);
    mix((uint64_t)B); mix(seed); mix(sample_id0); mix((uint64_t)loss_kind);
    if (!ctx->gexec || ctx->gkey != key) {
        if (ctx->gexec) { cudaGraphExecDestroy(ctx->gexec); ctx->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        RB_CUDA(cudaStreamSynchronize(st));
        RB_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeRelaxed));
        const int rc = issue(ctx->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        const cudaError_t ci = cudaGraphInstantiate(&ctx->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) { ctx->gexec = nullptr; return cuda_fail(ci, "cudaGraphInstantiate"); }
        ctx->gkey = key;
    }
    RB_CUDA(cudaGraphLaunch(ctx->gexec, st));
    return RB_OK;
}

}  // extern "C"
