// Generic dense stack on the tensor-core GEMM: the auxiliary objectives of the reference's backup experiments.
//   * value-function regressor  (/root/reference src/distilation/backup/student_rollout_mlp_vf.py:251-276): [prev_ob | next_ac] (13) -> 64 (linear)
//     -> 10 x tanh(100) -> 1, loss = sum (vpred - vtarg)^2, Adam lr 1e-2 (:285-295); targets from add_vtarg (:608-616)
//   * reward-prediction regressor (backup/student_rollout.py:161-164,328: dense 64 tanh -> 1 on the trunk features, loss += sum (reward - target)^2)
//   * any KL-trained dense student (loss.py:3-13 / backup/student_rollout.py:639-642) whose widths are not one of the two fused kernels
// Forward, loss and backward are calls of gemm_bf16x3 (bias + tanh fused in the forward epilogue, tanh' fused in the dgrad epilogue, split-K
// wgrad) plus fixed-order column sums and loss reductions, so results are bit-reproducible run to run.
// Flat parameter layout: for l = 1..L: W_l[d_{l-1}][d_l] (row-major, as tf.layers.dense kernels), b_l[d_l].
#include "common.cuh"
#include "gemm_tc.cuh"
#include "loss.cuh"

namespace rb {

constexpr int DENSE_MAX_LAYERS = RB_DENSE_MAX_LAYERS;

struct DenseSpec {
    int L;
    int d[DENSE_MAX_LAYERS + 1];
    int act[DENSE_MAX_LAYERS];
    int64_t w_off[DENSE_MAX_LAYERS], b_off[DENSE_MAX_LAYERS], P;
    size_t splitk_floats;
};

static int dense_parse(int L, const int* dims, const int* acts, DenseSpec& s) {
    RB_REQUIRE(dims != nullptr, "dims is NULL");
    RB_REQUIRE(L >= 1 && L <= DENSE_MAX_LAYERS, "layer count out of range");
    s.L = L; s.P = 0; s.splitk_floats = 0;
    for (int l = 0; l <= L; ++l) { RB_REQUIRE(dims[l] >= 1 && dims[l] <= 4096, "layer width out of range"); s.d[l] = dims[l]; }
    for (int l = 0; l < L; ++l) {
        s.act[l] = acts ? acts[l] : (l + 1 < L ? 1 : 0);
        RB_REQUIRE(s.act[l] == 0 || s.act[l] == 1, "activation must be 0 (linear) or 1 (tanh)");
        s.w_off[l] = s.P; s.P += (int64_t)s.d[l] * s.d[l + 1];
        s.b_off[l] = s.P; s.P += s.d[l + 1];
        s.splitk_floats = max(s.splitk_floats, (size_t)32 * s.d[l] * s.d[l + 1]);
    }
    return RB_OK;
}

// workspace (floats): a[1..L] and da[1..L] ([B, d_l] each), split-K partials, 1024 loss partials, column-sum scratch
struct DenseWs { float* a[DENSE_MAX_LAYERS + 1]; float* da[DENSE_MAX_LAYERS + 1]; float *splitk, *loss_part, *colpart; };
static size_t dense_ws_floats(const DenseSpec& s, int64_t B) {
    size_t n = 0;
    for (int l = 1; l <= s.L; ++l) n += 2 * (((size_t)B * s.d[l] + 3) & ~(size_t)3);
    return n + s.splitk_floats + 1024 + COLPART_FLOATS + 64;
}
static void dense_ws_carve(const DenseSpec& s, int64_t B, float* ws, DenseWs& w) {
    float* p = ws;
    auto take = [&](size_t n) { float* q = p; p += (n + 3) & ~(size_t)3; return q; };
    for (int l = 1; l <= s.L; ++l) { w.a[l] = take((size_t)B * s.d[l]); w.da[l] = take((size_t)B * s.d[l]); }
    w.splitk = take(s.splitk_floats); w.loss_part = take(1024); w.colpart = take(COLPART_FLOATS);
}

// loss and dL/ds over [B, n] outputs; grid-stride with a fixed grid => fixed summation order
__global__ void k_dense_loss(int64_t B, int n, const float* __restrict__ s, const float* __restrict__ t, int kind, float* __restrict__ ds,
                             float* __restrict__ loss_partial) {
    __shared__ float red[8];
    float l = 0.f;
    if (kind == RB_LOSS_MSE) {
        const int64_t total = B * n;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const float e = s[i] - t[i];
            l += e * e;
            ds[i] = 2.f * e;
        }
    } else {                                          // KL over pdflat rows (mean0, mean1, logstd0, logstd1): n == 4
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
            float4 d;
            l += kl_row(reinterpret_cast<const float4*>(s)[i], reinterpret_cast<const float4*>(t)[i], kind, d);
            reinterpret_cast<float4*>(ds)[i] = d;
        }
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

// add_vtarg (backup/student_rollout_mlp_vf.py:608-616), one thread per episode: target[T-1] = gamma^T r[T-1] (the reference's exponent),
// target[i] = gamma^i r[i] + target[i+1] for i = T-2 .. 0 -- an absolute-time discount, not a per-state return
__global__ void k_vf_targets(const float* __restrict__ rew, int64_t E, int T, float gamma, float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float* r = rew + e * T;
    float* o = out + e * T;
    float acc = powf(gamma, (float)T) * r[T - 1];
    o[T - 1] = acc;
    for (int i = T - 2; i >= 0; --i) {
        acc = fmaf(powf(gamma, (float)i), r[i], acc);
        o[i] = acc;
    }
}

#define RB_TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

static int dense_gemm(const float* A, int lda, int a_mn, const float* Bm, int ldb, int b_mn, float* C, int ldc, int M, int N, int K, const float* bias,
                      int act, const float* H, int ldh, float* splitk, size_t splitk_floats, int sms, cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = Bm; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.act = act; g.H = H; g.ldh = ldh; g.batch = 1;
    return gemm_bf16x3(g, splitk, splitk_floats, sms, st);
}

static int dense_run(const DenseSpec& s, const float* P, const float* x, const float* target, int64_t B, int loss_kind, bool fwd_only, float* out,
                     float* gradloss, float* ws, cudaStream_t st) {
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DenseWs w;
    dense_ws_carve(s, B, ws, w);
    const int Bi = (int)B, L = s.L;
    // ---- forward: a_l = act_l(a_{l-1} W_l + b_l) ---------------------------------------------------------------------------------
    const float* in = x;
    for (int l = 1; l <= L; ++l) {
        float* o = (l == L && out) ? out : w.a[l];
        RB_TRY(dense_gemm(in, s.d[l - 1], 0, P + s.w_off[l - 1], s.d[l], 1, o, s.d[l], Bi, s.d[l], s.d[l - 1], P + s.b_off[l - 1], s.act[l - 1], nullptr, 0,
                          nullptr, 0, sms, st));
        in = o;
    }
    if (fwd_only) return RB_OK;
    const float* sL = in;
    // ---- loss, dL/d(output) ---------------------------------------------------------------------------------------------------------
    const int64_t items = loss_kind == RB_LOSS_MSE ? B * s.d[L] : B;
    const unsigned blocks = (unsigned)min((int64_t)1024, (items + 255) / 256);
    k_dense_loss<<<blocks, 256, 0, st>>>(B, s.d[L], sL, target, loss_kind, w.da[L], w.loss_part);
    RB_CUDA(cudaGetLastError());
    RB_TRY(sum_serial(w.loss_part, (int)blocks, gradloss + s.P, st));
    // a linear output layer passes dL/ds straight through; a tanh output needs dL/dz = dL/ds (1 - s^2)
    RB_REQUIRE(s.act[L - 1] == 0, "the output layer must be linear");
    // ---- backward -------------------------------------------------------------------------------------------------------------------
    for (int l = L; l >= 1; --l) {
        const float* a_prev = l == 1 ? x : w.a[l - 1];
        const float* dz = w.da[l];
        float* gW = gradloss + s.w_off[l - 1];
        RB_TRY(dense_gemm(a_prev, s.d[l - 1], 1, dz, s.d[l], 1, gW, s.d[l], s.d[l - 1], s.d[l], Bi, nullptr, 0, nullptr, 0, w.splitk, s.splitk_floats, sms,
                          st));                                                                              // dW_l = a_{l-1}^T dz_l
        RB_TRY(colsum(dz, s.d[l], B, s.d[l], 1, 0, gradloss + s.b_off[l - 1], 0, w.colpart, st));           // db_l
        if (l > 1) {                                                                                         // dz_{l-1} = (dz_l W_l^T) * act'_{l-1}
            const bool th = s.act[l - 2] == 1;
            RB_TRY(dense_gemm(dz, s.d[l], 0, P + s.w_off[l - 1], s.d[l], 0, w.da[l - 1], s.d[l - 1], Bi, s.d[l - 1], s.d[l], nullptr, 0,
                              th ? w.a[l - 1] : nullptr, s.d[l - 1], nullptr, 0, sms, st));
        }
    }
    return RB_OK;
}

}  // namespace rb

using namespace rb;

extern "C" {

int64_t rb_dense_param_count(int n_layers, const int* dims) {
    DenseSpec s;
    return dense_parse(n_layers, dims, nullptr, s) ? -1 : s.P;
}

int64_t rb_dense_workspace_bytes(int n_layers, const int* dims, int64_t batch) {
    DenseSpec s;
    if (batch <= 0 || dense_parse(n_layers, dims, nullptr, s)) return -1;
    return (int64_t)(sizeof(float) * dense_ws_floats(s, batch));
}

int rb_dense_fwd(const float* params, int n_layers, const int* dims, const int* acts, const float* x, int64_t B, float* out, void* workspace,
                 void* stream) {
    RB_REQUIRE(params && x && out && workspace, "NULL argument");
    RB_REQUIRE(B > 0 && B < ((int64_t)1 << 30), "bad batch");
    DenseSpec s;
    int rc = dense_parse(n_layers, dims, acts, s);
    if (rc) return rc;
    return dense_run(s, params, x, nullptr, B, 0, true, out, nullptr, (float*)workspace, (cudaStream_t)stream);
}

int rb_dense_loss_grad(const float* params, int n_layers, const int* dims, const int* acts, const float* x, const float* target, int64_t B,
                       int loss_kind, float* out, float* gradloss, void* workspace, void* stream) {
    RB_REQUIRE(params && x && target && gradloss && workspace, "NULL argument");
    RB_REQUIRE(B > 0 && B < ((int64_t)1 << 30), "bad batch");
    RB_REQUIRE(loss_kind == RB_LOSS_KL_ST || loss_kind == RB_LOSS_KL_TS || loss_kind == RB_LOSS_MSE, "unknown loss kind");
    DenseSpec s;
    int rc = dense_parse(n_layers, dims, acts, s);
    if (rc) return rc;
    RB_REQUIRE(loss_kind == RB_LOSS_MSE || s.d[s.L] == 4, "the KL losses need a 4-wide pdflat output");
    return dense_run(s, params, x, target, B, loss_kind, false, out, gradloss, (float*)workspace, (cudaStream_t)stream);
}

int rb_vf_targets(const float* rew, int64_t episodes, int steps, float gamma, float* vtarg, void* stream) {
    RB_REQUIRE(rew && vtarg, "NULL argument");
    RB_REQUIRE(episodes > 0 && steps > 0, "bad shape");
    k_vf_targets<<<(unsigned)((episodes + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rew, episodes, steps, gamma, vtarg);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // extern "C"
