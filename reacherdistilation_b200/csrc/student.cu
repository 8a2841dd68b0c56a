// Student networks: forward, fused KL loss + backward (flat gradient), TF-form Adam -- RB_MODE_FP32 (CUDA-core) path.
// Replaces, in /root/reference:
//   student_mlp_graph                  src/distilation/student_nn.py:51-57      (RB_STUDENT_MLP   16-24-128-128-32-4)
//   StudentAgent / MlpPolicy 2x64      src/distilation/backup/student_rollout.py:79-87 (RB_STUDENT_POLICY64 11-64-64-4)
//   kl_loss                            src/distilation/loss.py:3-13 ; pd.kl both directions backup/student_rollout.py:639-642
//   sess.run([loss, minimize_adam])    src/distilation/mlp_train.py:148-161 ; lossandgrad backup/student_rollout.py:646,708
//   adam.minimize / MpiAdam.update     src/distilation/mlp_train.py:75-80 ; backup/student_rollout.py:658,709
//
// Kernel shape: persistent blocks, each looping over tiles of TILE samples.  Activations of every layer live in shared
// memory as [feature][sample] (sample contiguous, row stride TILE+4 floats) so that per-sample phases (forward, dgrad)
// and per-parameter phases (wgrad) are both bank-conflict free.  Each block accumulates its partial gradient in shared
// memory (one owner thread per parameter -> no atomics) and writes it to partials[block]; a second kernel sums the
// partials in block order, so the result is bit-reproducible run to run.
#include "common.cuh"
#include "loss.cuh"
#include "philox.cuh"

namespace rb {

constexpr int MAX_LAYERS = 5;
struct NetSpec {
    int nl;
    int dims[MAX_LAYERS + 1];
    int act[MAX_LAYERS];     // 1 = tanh
    int w_off[MAX_LAYERS];   // offsets into the flat parameter vector
    int b_off[MAX_LAYERS];
    int P;                   // parameter count
    int obfilter;            // 1: input z = clip((x - p[0:11]) / p[11:22], +-5)   (MlpPolicy)
    int smem_shift;          // floats of padding so that every w_off + shift is a multiple of 4 in the smem copy
    int rows;                // sum(dims)
};

static NetSpec make_spec(int kind) {
    NetSpec s{};
    if (kind == RB_STUDENT_POLICY64) {
        const PolicyOffsets o = policy_offsets(4);
        s.nl = 3;
        const int d[4] = {11, 64, 64, 4};
        for (int i = 0; i < 4; ++i) s.dims[i] = d[i];
        s.act[0] = 1; s.act[1] = 1; s.act[2] = 0;
        s.w_off[0] = o.W1; s.b_off[0] = o.b1; s.w_off[1] = o.W2; s.b_off[1] = o.b2; s.w_off[2] = o.W3; s.b_off[2] = o.b3;
        s.P = o.total; s.obfilter = 1; s.smem_shift = 2;
    } else {
        s.nl = 5;
        const int d[6] = {16, 24, 128, 128, 32, 4};
        for (int i = 0; i < 6; ++i) s.dims[i] = d[i];
        const int a[5] = {1, 1, 0, 1, 0};
        int off = 0;
        for (int i = 0; i < 5; ++i) {
            s.act[i] = a[i];
            s.w_off[i] = off; off += d[i] * d[i + 1];
            s.b_off[i] = off; off += d[i + 1];
        }
        s.P = off; s.obfilter = 0; s.smem_shift = 0;
    }
    s.rows = 0;
    for (int i = 0; i <= s.nl; ++i) s.rows += s.dims[i];
    return s;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// TILE samples per tile, TPS threads cooperate on one sample in the per-sample phases (blockDim = TILE * TPS).
template <int TILE, int TPS, bool W_IN_SMEM>
__global__ void __launch_bounds__(TILE* TPS) k_student(const NetSpec S, const float* __restrict__ params, const float* __restrict__ x,
                                                       const float* __restrict__ tpd, int64_t B, int loss_kind, int fwd_only,
                                                       float4* __restrict__ s_out, float* __restrict__ partials) {
    constexpr int LD = TILE + 4;
    constexpr int NT = TILE * TPS;
    extern __shared__ __align__(16) float smem[];
    float* acts = smem;                                   // [S.rows + 4][LD]  (+4 rows: teacher pdflat)
    float* gacc = acts + (size_t)(S.rows + 4) * LD;       // [P + 4] gradient accumulators (+ loss at P), fwd_only: unused
    float* wsm = gacc + (fwd_only ? 0 : ((S.P + 4 + 3) & ~3));   // optional aligned copy of the parameters
    __shared__ float red[NT / 32];

    const int tid = threadIdx.x;
    const int s = tid % TILE, part = tid / TILE;
    const float* Wp = params;   // generic pointer: parameters in global, or the shifted smem copy
    if (W_IN_SMEM) {
        for (int i = tid; i < S.P; i += NT) wsm[i + S.smem_shift] = __ldg(params + i);
        Wp = wsm + S.smem_shift;
    }
    if (!fwd_only) for (int i = tid; i < S.P + 4; i += NT) gacc[i] = 0.f;
    __syncthreads();

    int row_of[MAX_LAYERS + 1];
    row_of[0] = 0;
#pragma unroll
    for (int l = 0; l < MAX_LAYERS; ++l) row_of[l + 1] = row_of[l] + (l <= S.nl ? S.dims[l] : 0);
    float* trow = acts + (size_t)S.rows * LD;
    const int in0 = S.dims[0];
    const int64_t ntiles = (B + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int nvalid = (int)min((int64_t)TILE, B - base);
        // ---- load inputs (transpose to [feature][sample]); rows past B are zero ---------------------------------
        for (int idx = tid; idx < TILE * in0; idx += NT) {
            const int ss = idx / in0, k = idx - ss * in0;
            float v = ss < nvalid ? __ldg(x + base * in0 + idx) : 0.f;
            if (S.obfilter) v = fminf(5.f, fmaxf(-5.f, (v - Wp[k]) / Wp[11 + k]));
            acts[k * LD + ss] = v;
        }
        if (!fwd_only)
            for (int idx = tid; idx < TILE * 4; idx += NT) {
                const int ss = idx >> 2, k = idx & 3;
                trow[k * LD + ss] = ss < nvalid ? __ldg(tpd + base * 4 + idx) : 0.f;
            }
        __syncthreads();
        // ---- forward -----------------------------------------------------------------------------------------
        for (int l = 0; l < S.nl; ++l) {
            const int in = S.dims[l], out = S.dims[l + 1];
            const float* X = acts + (size_t)row_of[l] * LD;
            float* Y = acts + (size_t)row_of[l + 1] * LD;
            const float* W = Wp + S.w_off[l];
            const float* bb = Wp + S.b_off[l];
            for (int j0 = part * 4; j0 < out; j0 += 4 * TPS) {
                float4 acc = *reinterpret_cast<const float4*>(bb + j0);
                for (int k = 0; k < in; ++k) {
                    const float a = X[k * LD + s];
                    const float4 w = *reinterpret_cast<const float4*>(W + k * out + j0);
                    acc.x = fmaf(a, w.x, acc.x); acc.y = fmaf(a, w.y, acc.y); acc.z = fmaf(a, w.z, acc.z); acc.w = fmaf(a, w.w, acc.w);
                }
                if (S.act[l]) { acc.x = tanhf(acc.x); acc.y = tanhf(acc.y); acc.z = tanhf(acc.z); acc.w = tanhf(acc.w); }
                Y[(j0 + 0) * LD + s] = acc.x; Y[(j0 + 1) * LD + s] = acc.y; Y[(j0 + 2) * LD + s] = acc.z; Y[(j0 + 3) * LD + s] = acc.w;
            }
            __syncthreads();
        }
        // ---- output, loss and dL/dpdflat (in place over the output rows) ----------------------------------------
        float* O = acts + (size_t)row_of[S.nl] * LD;
        float lsum = 0.f;
        if (part == 0) {
            const float m0 = O[0 * LD + s], m1 = O[1 * LD + s], l0 = O[2 * LD + s], l1 = O[3 * LD + s];
            if (s < nvalid && s_out) s_out[base + s] = make_float4(m0, m1, l0, l1);
            if (!fwd_only) {
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                if (s < nvalid) {
                    float4 dd;
                    lsum = pd_loss_row(make_float4(m0, m1, l0, l1), make_float4(trow[0 * LD + s], trow[1 * LD + s], trow[2 * LD + s], trow[3 * LD + s]),
                                       loss_kind, dd);
                    d0 = dd.x; d1 = dd.y; d2 = dd.z; d3 = dd.w;
                }
                O[0 * LD + s] = d0; O[1 * LD + s] = d1; O[2 * LD + s] = d2; O[3 * LD + s] = d3;
            }
        }
        if (fwd_only) { __syncthreads(); continue; }
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) red[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; ++w) t += red[w];
            gacc[S.P] += t;
        }
        // ---- backward ----------------------------------------------------------------------------------------
        for (int l = S.nl - 1; l >= 0; --l) {
            const int in = S.dims[l], out = S.dims[l + 1];
            float* X = acts + (size_t)row_of[l] * LD;          // layer input (post-activation of layer l-1)
            const float* D = acts + (size_t)row_of[l + 1] * LD; // dL/d(pre-activation of layer l)
            const int oq = out >> 2;
            // wgrad: one owner thread per (k, 4 consecutive j); k == in is the bias row
            for (int q = tid; q < (in + 1) * oq; q += NT) {
                const int k = q / oq, j0 = (q - k * oq) << 2;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* d0 = D + (size_t)j0 * LD;
                if (k < in) {
                    const float* xr = X + (size_t)k * LD;
                    for (int ss = 0; ss < TILE; ss += 4) {
                        const float4 xv = *reinterpret_cast<const float4*>(xr + ss);
                        const float4 a = *reinterpret_cast<const float4*>(d0 + ss);
                        const float4 b = *reinterpret_cast<const float4*>(d0 + LD + ss);
                        const float4 c = *reinterpret_cast<const float4*>(d0 + 2 * LD + ss);
                        const float4 d = *reinterpret_cast<const float4*>(d0 + 3 * LD + ss);
                        acc.x = fmaf(xv.x, a.x, fmaf(xv.y, a.y, fmaf(xv.z, a.z, fmaf(xv.w, a.w, acc.x))));
                        acc.y = fmaf(xv.x, b.x, fmaf(xv.y, b.y, fmaf(xv.z, b.z, fmaf(xv.w, b.w, acc.y))));
                        acc.z = fmaf(xv.x, c.x, fmaf(xv.y, c.y, fmaf(xv.z, c.z, fmaf(xv.w, c.w, acc.z))));
                        acc.w = fmaf(xv.x, d.x, fmaf(xv.y, d.y, fmaf(xv.z, d.z, fmaf(xv.w, d.w, acc.w))));
                    }
                    float* g = gacc + S.w_off[l] + k * out + j0;
                    g[0] += acc.x; g[1] += acc.y; g[2] += acc.z; g[3] += acc.w;
                } else {
                    for (int ss = 0; ss < TILE; ss += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(d0 + ss);
                        const float4 b = *reinterpret_cast<const float4*>(d0 + LD + ss);
                        const float4 c = *reinterpret_cast<const float4*>(d0 + 2 * LD + ss);
                        const float4 d = *reinterpret_cast<const float4*>(d0 + 3 * LD + ss);
                        acc.x += (a.x + a.y) + (a.z + a.w); acc.y += (b.x + b.y) + (b.z + b.w);
                        acc.z += (c.x + c.y) + (c.z + c.w); acc.w += (d.x + d.y) + (d.z + d.w);
                    }
                    float* g = gacc + S.b_off[l] + j0;
                    g[0] += acc.x; g[1] += acc.y; g[2] += acc.z; g[3] += acc.w;
                }
            }
            __syncthreads();
            if (l == 0) break;
            // dgrad into the input rows (in place): dX[k] = (sum_j D[j] W[k][j]) * tanh'(X[k])
            const float* W = Wp + S.w_off[l];
            for (int k = part; k < in; k += TPS) {
                float a0 = 0.f, a1 = 0.f;
                const float* wr = W + (size_t)k * out;
                for (int j = 0; j < out; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + j);
                    a0 = fmaf(D[(j + 0) * LD + s], w.x, a0); a1 = fmaf(D[(j + 1) * LD + s], w.y, a1);
                    a0 = fmaf(D[(j + 2) * LD + s], w.z, a0); a1 = fmaf(D[(j + 3) * LD + s], w.w, a1);
                }
                float d = a0 + a1;
                if (S.act[l - 1]) { const float h = X[k * LD + s]; d *= (1.f - h * h); }
                X[k * LD + s] = d;
            }
            __syncthreads();
        }
    }
    if (!fwd_only) {
        __syncthreads();
        float* out = partials + (size_t)blockIdx.x * (S.P + 1);
        for (int i = tid; i < S.P + 1; i += NT) out[i] = gacc[i];
    }
}

__global__ void k_reduce_partials(const float* __restrict__ partials, int nblocks, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += partials[(size_t)b * n + i];
    out[i] = acc;
}

// TF1 Adam ("epsilon hat"): mlp_train.py:75-80; MpiAdam.update backup/student_rollout.py:709
__global__ void k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, int64_t n,
                       float lr_t, float b1, float b2, float eps, float gscale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g[i], lr_t, b1, b2, eps, gscale);
    p[i] = pi; m[i] = mi; v[i] = vi;
}

// mlp_train.py:50-52: x = concat(dropout(ob, kp), prev_pdflat, prev_rew)
__global__ void k_mlp_input(const float* __restrict__ obs, const float4* __restrict__ prev_pd, const float* __restrict__ prev_rew,
                            int64_t B, float keep_prob, uint32_t k0, uint32_t k1, uint32_t sample_id0, uint32_t iteration,
                            float4* __restrict__ x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float ob[12];
#pragma unroll
    for (int k = 0; k < 11; ++k) ob[k] = __ldg(obs + i * 11 + k);
    ob[11] = 0.f;
    if (keep_prob < 1.f) {
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            const uint4 r = philox4x32_10(sample_id0 + (uint32_t)i, iteration, (uint32_t)blk, STREAM_DROPOUT, k0, k1);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float u = (float)(rr[c] >> 8) * 5.9604644775390625e-08f;
                const float keep = floorf(keep_prob + u);
                ob[4 * blk + c] = __fdiv_rn(ob[4 * blk + c], keep_prob) * keep;
            }
        }
    }
    const float4 pp = prev_pd ? __ldg(prev_pd + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float pr = prev_rew ? __ldg(prev_rew + i) : 0.f;
    x[i * 4 + 0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
    x[i * 4 + 1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
    x[i * 4 + 2] = make_float4(ob[8], ob[9], ob[10], pp.x);
    x[i * 4 + 3] = make_float4(pp.y, pp.z, pp.w, pr);
}

static int g_sm_count[16] = {0};
static int sm_count_of(int device) {
    if (device < 0 || device >= 16) return 148;
    if (!g_sm_count[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
        g_sm_count[device] = v;
    }
    return g_sm_count[device];
}
constexpr int MAX_STUDENT_BLOCKS = 2 * 160;

template <int TILE, int TPS, bool WS>
static int launch_student(const NetSpec& S, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind,
                          int fwd_only, float* s_out, float* gradloss, float* partials, cudaStream_t st) {
    int device = 0;
    RB_CUDA(cudaGetDevice(&device));
    const int64_t ntiles = (B + TILE - 1) / TILE;
    size_t smem = sizeof(float) * ((size_t)(S.rows + 4) * (TILE + 4) + (fwd_only ? 0 : ((S.P + 4 + 3) & ~3)) + (WS ? (S.P + 8) : 0));
    const int per_sm = smem * 2 + 4096 <= 227 * 1024 ? 2 : 1;
    const int grid = (int)min((int64_t)min(per_sm * sm_count_of(device), MAX_STUDENT_BLOCKS), ntiles);
    auto kern = k_student<TILE, TPS, WS>;
    static bool seen[RB_MAX_DEVICES] = {};             // function attributes are per device
    if (first_use_on_device(seen)) RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    RB_REQUIRE(smem <= 224 * 1024, "shared memory budget exceeded");
    RB_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "params / x must be 16-byte aligned");
    kern<<<grid, TILE * TPS, smem, st>>>(S, params, x, tpd, B, loss_kind, fwd_only, (float4*)s_out, partials);
    RB_CUDA(cudaGetLastError());
    if (!fwd_only) {
        const int n = S.P + 1;
        k_reduce_partials<<<(n + 255) / 256, 256, 0, st>>>(partials, grid, n, gradloss);
        RB_CUDA(cudaGetLastError());
    }
    return RB_OK;
}

// implemented in student_tc.cu (tcgen05 path)
struct AdamFuse { float* p; float* m; float* v; float lr_t, beta1, beta2, eps, gscale; };
struct PeerExchange { int world, rank; uint32_t epoch; const uint64_t* gl_ptrs; const uint64_t* flag_ptrs; const uint64_t* gl_ptrs_alt; };
struct StepClock { const uint32_t* clock; float lr, beta1, beta2; };
int student_tc_run(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only, float* s_out,
                   float* gradloss, void* workspace, const AdamFuse* adam, const PeerExchange* px, const StepClock* clk, cudaStream_t st);
size_t student_tc_workspace_floats();

}  // namespace rb

using namespace rb;

extern "C" {

int64_t rb_student_param_count(int kind) {
    if (kind != RB_STUDENT_POLICY64 && kind != RB_STUDENT_MLP) return -1;
    return make_spec(kind).P;
}
int rb_student_input_dim(int kind) {
    if (kind != RB_STUDENT_POLICY64 && kind != RB_STUDENT_MLP) return -1;
    return make_spec(kind).dims[0];
}
int64_t rb_student_workspace_bytes(int kind, int64_t batch, int device) {
    (void)batch; (void)device;
    if (kind != RB_STUDENT_POLICY64 && kind != RB_STUDENT_MLP) return -1;
    const size_t fp32_floats = (size_t)MAX_STUDENT_BLOCKS * (make_spec(kind).P + 1);
    return (int64_t)(sizeof(float) * max(fp32_floats, student_tc_workspace_floats()));
}

int rb_student_mlp_input(const float* obs, const float* prev_pd, const float* prev_rew, int64_t B, float keep_prob, uint64_t seed,
                         uint32_t sample_id0, uint32_t iteration, float* x, void* stream) {
    RB_REQUIRE(obs && x, "NULL argument");
    RB_REQUIRE(keep_prob > 0.f, "keep_prob must be > 0");
    if (B <= 0) return RB_OK;
    k_mlp_input<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(obs, (const float4*)prev_pd, prev_rew, B, keep_prob,
                                                                              (uint32_t)seed, (uint32_t)(seed >> 32), sample_id0, iteration,
                                                                              (float4*)x);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

static int student_dispatch(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only,
                            float* s_out, float* gradloss, void* ws, cudaStream_t st) {
    const NetSpec S = make_spec(kind);
    if (kind == RB_STUDENT_POLICY64) return launch_student<128, 2, true>(S, params, x, tpd, B, loss_kind, fwd_only, s_out, gradloss, (float*)ws, st);
    return launch_student<64, 4, false>(S, params, x, tpd, B, loss_kind, fwd_only, s_out, gradloss, (float*)ws, st);
}

int rb_student_fwd_ws(int kind, const float* params, const float* x, int64_t B, float* s_pd, void* ws, int mode, void* stream) {
    RB_REQUIRE(params && x && s_pd, "NULL argument");
    RB_REQUIRE(kind == RB_STUDENT_POLICY64 || kind == RB_STUDENT_MLP, "unknown student kind");
    if (B <= 0) return RB_OK;
    if (mode == RB_MODE_TC) return student_tc_run(kind, params, x, nullptr, B, 0, 1, s_pd, nullptr, ws, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    RB_REQUIRE(mode == RB_MODE_FP32, "unknown mode");
    return student_dispatch(kind, params, x, nullptr, B, 0, 1, s_pd, nullptr, nullptr, (cudaStream_t)stream);
}

int rb_student_fwd(int kind, const float* params, const float* x, int64_t B, float* s_pd, int mode, void* stream) {
    RB_REQUIRE(params && x && s_pd, "NULL argument");
    RB_REQUIRE(kind == RB_STUDENT_POLICY64 || kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(mode == RB_MODE_FP32, "rb_student_fwd: RB_MODE_TC needs a workspace, use rb_student_fwd_ws");
    if (B <= 0) return RB_OK;
    return student_dispatch(kind, params, x, nullptr, B, 0, 1, s_pd, nullptr, nullptr, (cudaStream_t)stream);
}

int rb_student_loss_grad(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, float* s_pd,
                         float* gradloss, void* ws, int mode, void* stream) {
    RB_REQUIRE(params && x && tpd && gradloss && ws, "NULL argument");
    RB_REQUIRE(kind == RB_STUDENT_POLICY64 || kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    RB_REQUIRE(B > 0, "empty batch");
    if (mode == RB_MODE_TC) return student_tc_run(kind, params, x, tpd, B, loss_kind, 0, s_pd, gradloss, ws, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    RB_REQUIRE(mode == RB_MODE_FP32, "unknown mode");
    return student_dispatch(kind, params, x, tpd, B, loss_kind, 0, s_pd, gradloss, ws, (cudaStream_t)stream);
}

static float adam_lr_t(float lr, float b1, float b2, int64_t t) {
    return (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}

int rb_student_step(int kind, float* params, float* m, float* v, const float* x, const float* tpd, int64_t B, int loss_kind, float* s_pd,
                    float* gradloss, void* ws, int64_t t, float lr, float b1, float b2, float eps, float gscale, int mode, void* stream) {
    RB_REQUIRE(params && m && v && x && tpd && gradloss && ws, "NULL argument");
    RB_REQUIRE(kind == RB_STUDENT_POLICY64 || kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    RB_REQUIRE(B > 0 && t >= 1, "empty batch / bad step");
    if (mode == RB_MODE_TC) {
        const AdamFuse af{params, m, v, adam_lr_t(lr, b1, b2, t), b1, b2, eps, gscale};
        return student_tc_run(kind, params, x, tpd, B, loss_kind, 0, s_pd, gradloss, ws, &af, nullptr, nullptr, (cudaStream_t)stream);
    }
    int rc = rb_student_loss_grad(kind, params, x, tpd, B, loss_kind, s_pd, gradloss, ws, mode, stream);
    if (rc) return rc;
    return rb_adam_step(params, m, v, gradloss, rb_student_param_count(kind), t, lr, b1, b2, eps, gscale, stream);
}

int rb_student_step_dp(int kind, float* params, float* m, float* v, const float* x, const float* tpd, int64_t B, int loss_kind, float* s_pd,
                       float* gradloss, void* ws, int64_t t, float lr, float b1, float b2, float eps, float gscale, int rank, int world,
                       const uint64_t* peer_grad_slots, const uint64_t* peer_flags, uint32_t epoch, void* stream) {
    RB_REQUIRE(params && m && v && x && tpd && gradloss && ws && peer_grad_slots && peer_flags, "NULL argument");
    RB_REQUIRE(kind == RB_STUDENT_POLICY64 || kind == RB_STUDENT_MLP, "unknown student kind");
    RB_REQUIRE(pd_loss_kind_ok(loss_kind), "unknown loss kind");
    RB_REQUIRE(B > 0 && t >= 1 && epoch >= 1, "empty batch / bad step / epoch must start at 1");
    RB_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, "2..8 ranks");
    const AdamFuse af{params, m, v, adam_lr_t(lr, b1, b2, t), b1, b2, eps, gscale};
    const PeerExchange px{world, rank, epoch, peer_grad_slots, peer_flags, nullptr};
    return student_tc_run(kind, params, x, tpd, B, loss_kind, 0, s_pd, gradloss, ws, &af, &px, nullptr, (cudaStream_t)stream);
}

int rb_adam_step(float* p, float* m, float* v, const float* g, int64_t P, int64_t t, float lr, float b1, float b2, float eps,
                 float gscale, void* stream) {
    RB_REQUIRE(p && m && v && g, "NULL argument");
    RB_REQUIRE(P > 0 && t >= 1, "bad size / step");
    const double lr_t = adam_lr_t(lr, b1, b2, t);
    k_adam<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, m, v, g, P, (float)lr_t, b1, b2, eps, gscale);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // extern "C"
