// RB_MODE_TC student: forward, fused KL loss, backward (dgrad + wgrad) on the 5th-gen tensor cores.
// Replaces sess.run([loss, minimize_adam]) minus Adam (/root/reference src/distilation/mlp_train.py:148-161), the student graph
// student_nn.py:51-57 (16-24-128-128(lin)-32-4), the backup student backup/student_rollout.py:79-87 (11-64-64-4, obfilter),
// kl_loss loss.py:3-13 and lossandgrad backup/student_rollout.py:639-646,708.
//
// One persistent CTA per SM (512 threads) walks over tiles of 128 samples.  Design:
//  * ACTIVATIONS live in one feature-major shared-memory matrix per bf16 split half (ACT_HI, ACT_LO): "group" g = 8 features
//    x 128 samples = 2048 B at byte g*2048, sample s at +s*16, feature f%8 at +(f%8)*2.  That is the UMMA no-swizzle K-major
//    layout (LBO 2048, SBO 128) when a window of groups is used as the A operand of forward / dgrad GEMMs (M = samples), and
//    the SAME BYTES are the MN-major layout (K-group stride 128, MN-group stride 2048) when a window is used as an operand of
//    the wgrad GEMMs (K = samples), so nothing is ever transposed or copied.
//  * Each layer input X_l is followed by a constant ONES group (feature 0 == 1): the forward GEMM picks the bias up from an
//    extra K row of the weight tile, and the wgrad GEMM  G_l[out, in|1] += dZ_l^T [X_l | 1]  yields the bias gradient as one
//    more column.  dZ_l (gradient w.r.t. the pre-activation of layer l) overwrites X_{l+1} in place.
//  * WEIGHTS: one K-major tile per layer (hi and lo), element (n = out, k = in); the dgrad GEMM reads the same tile through
//    an MN-major descriptor (dX = dZ W^T).
//  * ACCUMULATORS in TMEM: columns [0,128) forward / dgrad results; G_l of every layer behind them.  The weight-gradient
//    accumulators stay in TMEM across ALL tiles of the CTA and are written out once (fixed order => bit-reproducible).
//  * bf16x3 operand splitting (hi*hi + lo*hi + hi*lo, fp32 accumulate) keeps results ~1e-5 of fp32.
//  * The reference's third layer is LINEAR (student_nn.py:55), so layers 3 and 4 are folded for the GEMMs:
//    W34 = W3 W4, b34 = b3 W4 + b4 (k_fold34, fp32).  The tile kernel accumulates G34 = dL/dW34, g34 = dL/db34 and
//    k_student_finish maps them back exactly:  dW3 = G34 W4^T, db3 = W4 g34, dW4 = W3^T G34 + b3 (x) g34, db4 = g34.
//    This removes the 128x128 layer (60 % of the MMA work and 100 KB of shared memory) without changing the function.
#include <cooperative_groups.h>

#include "common.cuh"
#include "loss.cuh"
#include "physics.cuh"
#include "tc_common.cuh"

namespace rb {

using namespace tc;

constexpr int ST_THREADS = 512;
constexpr int ST_TILE = 128;
constexpr int ST_GROUPS = 32;                         // groups per ACT half
constexpr int ST_ACT_BYTES = ST_GROUPS * 2048;        // 64 KB
constexpr int ST_MAXL = 4;
constexpr int ST_MAX_GRID = 160;                      // cooperative grid = one CTA per SM, at most this many

// ---- network specs (the folded MLP and the 2x64 policy student) ---------------------------------------------------------
struct SpecMLP {   // 16 -> 24 tanh -> 128 tanh -> [128 lin -> ] 32 tanh -> 4
    static constexpr int L = 4, IN0 = 16, OBFILTER = 0;
    __host__ __device__ static constexpr int in(int l) { return l == 0 ? 16 : l == 1 ? 24 : l == 2 ? 128 : 32; }
    __host__ __device__ static constexpr int out(int l) { return l == 0 ? 24 : l == 1 ? 128 : l == 2 ? 32 : 4; }
    // group index of X_l in the ACT halves (ONES group follows each); dZ of the last layer sits at group 0
    __host__ __device__ static constexpr int slot(int l) { return l == 0 ? 7 : l == 1 ? 10 : l == 2 ? 14 : 2; }
};
struct SpecPOL {   // 11 -> 64 tanh -> 64 tanh -> 4, input z = clip((x - mu) / sd, +-5)
    static constexpr int L = 3, IN0 = 11, OBFILTER = 1;
    __host__ __device__ static constexpr int in(int l) { return l == 0 ? 11 : 64; }
    __host__ __device__ static constexpr int out(int l) { return l == 2 ? 4 : 64; }
    __host__ __device__ static constexpr int slot(int l) { return l == 0 ? 2 : l == 1 ? 5 : 14; }
};
template <class S> struct Geo {
    __host__ __device__ static constexpr int ig(int l) { return (S::in(l) + 7) / 8; }                  // feature groups of X_l
    __host__ __device__ static constexpr int K(int l) { return ((ig(l) * 8 + 1 + 15) / 16) * 16; }     // forward K incl. ones
    __host__ __device__ static constexpr int N(int l) { return ((S::out(l) + 15) / 16) * 16; }         // forward N
    __host__ __device__ static constexpr int inpad(int l) { return ((S::in(l) + 15) / 16) * 16; }      // dgrad N
    __host__ __device__ static constexpr int dzslot(int l) { return l == S::L - 1 ? 0 : S::slot(l + 1); }
    __host__ __device__ static constexpr int wtile_off(int l) { return l == 0 ? 0 : wtile_off(l - 1) + N(l - 1) * K(l - 1) * 2; }   // bytes
    __host__ __device__ static constexpr int wtile_bytes() { return wtile_off(S::L); }
    // weight-gradient accumulator G_l in TMEM.  "P": lanes = out features, columns = [in | 1] (K(l) columns).  "Q" (fewer columns when
    // the layer narrows): lanes = [in | 1] (the window at X_l), columns = out (N(l)); if the ONES row falls outside the 128 lanes
    // (in == 128) the bias gradient comes from one more small P-oriented product (16 columns).
    __host__ __device__ static constexpr bool wq(int l) { return N(l) < K(l); }
    __host__ __device__ static constexpr bool wq_extra(int l) { return wq(l) && ig(l) * 8 + 1 > 128; }
    __host__ __device__ static constexpr int gcols(int l) { return wq(l) ? N(l) + (wq_extra(l) ? 16 : 0) : K(l); }
    __host__ __device__ static constexpr int gcol(int l) { return l == 0 ? 128 : gcol(l - 1) + gcols(l - 1); }   // TMEM column of G_l
};

struct StudentTcArgs {
    const float* w[ST_MAXL];      // W_l row-major [in][out]
    const float* b[ST_MAXL];      // b_l [out]
    int pw[ST_MAXL], pb[ST_MAXL]; // offsets of dW_l / db_l inside one partial vector
    int ploss, pstride;           // offset of the loss, floats per partial vector
    const float* obf;             // SpecPOL: ob_mean[11] ob_std[11]
    const float* x;               // [B, IN0]
    const float* x_act;           // SpecMLP, optional: [B, 16] the same rows WITHOUT observation dropout.  When given, every tile runs the forward
                                  // pass twice: first on x_act -> s_out (the pdflat the student ACTS with: the reference acts with keep_prob 1,
                                  // mlp_train.py:171-186), then on x for the loss and the gradient (keep_prob = KEEP_PROB, mlp_train.py:151)
    const float* t;               // [B, 4] teacher pdflat (NULL: forward only)
    float4* s_out;                // [B] student pdflat (may be NULL)
    float* partials;              // [grid][pstride]
    int64_t B;
    int loss_kind, fwd_only;
    // cooperative phases (all global scratch lives in the caller's workspace)
    const float* params;          // flat parameter vector (fold / finish / Adam)
    float* ctlws;                 // [64] scalars handed from k_student_image to k_student_tc: [0] = Adam step size lr_t of this step (device-clock mode)
    float* snap;                  // SpecMLP: snapshot of W3 | b3 | W4 (params[M_W3 .. M_B4)) taken by k_student_image: what the un-fold reads
    uint8_t* wimg;                // split weight image: hi tiles then lo tiles, exactly the shared-memory layout (built by k_student_image)
    float* red;                   // reduced partial vector [pstride]
    float* gradloss;              // final flat gradient [P] + loss
    int P;                        // parameter count
    // optional fused TF-form Adam (single rank: no all-reduce between gradient and update)
    int do_adam;
    float* adam_p; float* adam_m; float* adam_v;
    float lr_t, beta1, beta2, eps, gscale;
    // optional data-parallel exchange fused in front of the update: one-shot all-reduce over NVLink peer memory.
    // peer_ll[r] = rank r's receive area of this step: uint2 {value bits, epoch} [world][SL], SL = round_up(P + 1, 64); this rank writes
    // row `rank` of every area (symmetric allocation, double buffered by the caller: peer_ll2[epoch & 1] when the step clock is used)
    int world, rank;
    uint32_t epoch;
    uint2* peer_ll[8];
    // When set, adam step = clock[1] + 1 (lr_t computed in-kernel from `lr`), epoch = clock[2] + 1, areas = peer_ll2[epoch & 1].
    const uint32_t* clock;
    float lr;
    uint2* peer_ll2[2][8];
    // optional fused DAgger env step (rb_dagger_step; sample i == env i): once the forward pass of a tile has written s_out, the envs of
    // that tile are stepped with the student mean INSIDE this launch by the CTAs that own one tile less than the others and would
    // otherwise idle at the first grid barrier (per-tile release / acquire flags): two 512-env blocks each, i.e. all envs at the config-4
    // shard (32 768 envs: 40 such CTAs).  Envs beyond that (student_tc_act_covered()) are left to a k_dagger_act launch behind this
    // kernel.  One thread of the grid advances the device clock and posts the loss mailbox at the end.
    int act_on;
    float4* act_qv; float4* act_tp; uint4* act_ctr;
    float4* act_prev_t; float* act_prev_rec_rew; float* act_last_reward; float* act_rew; uint8_t* act_done;
    uint32_t act_k0, act_k1, act_offset;
    uint32_t* act_flags;          // [ceil(B / 128)] forward-done flag per tile, value = iteration count of the launch that set it
    uint32_t* act_clock;          // writable alias of `clock`
    uint2* act_mailbox;           // mapped host memory {loss bits, iterations done}, may be NULL
};

// Plain (weak) global load.  Data produced earlier in the SAME launch by other CTAs is read only after a grid barrier and is never
// touched by this SM before that barrier, so no stale L1 line can exist; unlike ld.global.cg / volatile these loads may be batched.
template <typename T> __device__ __forceinline__ T ldw(const T* p) { return *p; }

// {value, epoch} pair as ONE 8-byte store / load at system scope (peer memory over NVLink; bypasses L1 on the polling side)
__device__ __forceinline__ void st_ll(uint2* p, uint32_t v, uint32_t e) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(e) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
    uint2 w;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

struct __align__(16) StudentTcCtl {
    uint32_t epoch;
    uint32_t iter;                // iterations done once this launch has finished (clock[0] + 1): value of the act flags / mailbox
    uint64_t mbar;                // forward layers / dgrad results
    uint64_t mbar2;               // wgrad completion (X_l may be overwritten)
    uint64_t mbar_load;           // TMA bulk copy of the weight image
    uint32_t tmem_base;
    float red[ST_THREADS / 32];
};

__device__ __forceinline__ uint32_t make_idesc_bf16_ex(int M, int N, int a_mn, int b_mn) {
    return make_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// three bf16x3 terms for one K-step
__device__ __forceinline__ void mma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    mma_bf16(d, ah, bh, idesc, acc);
    mma_bf16(d, al, bh, idesc, 1);
    mma_bf16(d, ah, bl, idesc, 1);
}

template <class S, int l> __device__ __forceinline__ void issue_fwd(uint32_t tmem, uint32_t act_hi, uint32_t act_lo, uint32_t w_hi, uint32_t w_lo) {
    using G = Geo<S>;
    constexpr int K = G::K(l), N = G::N(l);
    const uint32_t idesc = make_idesc_bf16_ex(128, N, 0, 0);
    const uint32_t a0 = S::slot(l) * 2048, b0 = G::wtile_off(l);
    uint64_t dah = make_smem_desc(act_hi + a0, 2048, 128), dal = make_smem_desc(act_lo + a0, 2048, 128);
    uint64_t dbh = make_smem_desc(w_hi + b0, N * 16, 128), dbl = make_smem_desc(w_lo + b0, N * 16, 128);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
        mma3(tmem, dah, dal, dbh, dbl, idesc, ks > 0);
        dah = desc_advance(dah, 2 * 2048); dal = desc_advance(dal, 2 * 2048);
        dbh = desc_advance(dbh, 2 * N * 16); dbl = desc_advance(dbl, 2 * N * 16);
    }
}
// dX_l[128, inpad] = dZ_l[128, N_l] * W_l^T : A = dZ window (K-major), B = weight tile through an MN-major descriptor
template <class S, int l> __device__ __forceinline__ void issue_dgrad(uint32_t tmem, uint32_t act_hi, uint32_t act_lo, uint32_t w_hi, uint32_t w_lo) {
    using G = Geo<S>;
    constexpr int N = G::N(l), NP = G::inpad(l);
    const uint32_t idesc = make_idesc_bf16_ex(128, NP, 0, 1);
    const uint32_t a0 = G::dzslot(l) * 2048, b0 = G::wtile_off(l);
    uint64_t dah = make_smem_desc(act_hi + a0, 2048, 128), dal = make_smem_desc(act_lo + a0, 2048, 128);
    uint64_t dbh = make_smem_desc(w_hi + b0, 128, N * 16), dbl = make_smem_desc(w_lo + b0, 128, N * 16);
#pragma unroll
    for (int ks = 0; ks < N / 16; ++ks) {
        mma3(tmem, dah, dal, dbh, dbl, idesc, ks > 0);
        dah = desc_advance(dah, 2 * 2048); dal = desc_advance(dal, 2 * 2048);
        dbh = desc_advance(dbh, 2 * 128); dbl = desc_advance(dbl, 2 * 128);
    }
}
// G_l += (see Geo::wq): both operands are MN-major windows of ACT, K = the 128 samples of the tile
template <class S, int l> __device__ __forceinline__ void issue_wgrad(uint32_t tmem, uint32_t act_hi, uint32_t act_lo, uint32_t first) {
    using G = Geo<S>;
    const uint32_t xs = S::slot(l) * 2048, zs = G::dzslot(l) * 2048, d = tmem + G::gcol(l);
    const uint32_t acc0 = first ? 0u : 1u;
    if constexpr (G::wq(l)) {
        const uint32_t idesc = make_idesc_bf16_ex(128, G::N(l), 1, 1);                 // D[in|1, out] += [X_l | 1]^T dZ_l
        uint64_t dah = make_smem_desc(act_hi + xs, 128, 2048), dal = make_smem_desc(act_lo + xs, 128, 2048);
        uint64_t dbh = make_smem_desc(act_hi + zs, 128, 2048), dbl = make_smem_desc(act_lo + zs, 128, 2048);
#pragma unroll
        for (int ks = 0; ks < ST_TILE / 16; ++ks) {
            mma3(d, dah, dal, dbh, dbl, idesc, ks > 0 ? 1u : acc0);
            dah = desc_advance(dah, 256); dal = desc_advance(dal, 256); dbh = desc_advance(dbh, 256); dbl = desc_advance(dbl, 256);
        }
        if constexpr (G::wq_extra(l)) {                                                // D2[out, 16] += dZ_l^T [1 | .]  (column 0 = bias gradient)
            const uint32_t idesc2 = make_idesc_bf16_ex(128, 16, 1, 1), os = (S::slot(l) + G::ig(l)) * 2048;
            uint64_t eah = make_smem_desc(act_hi + zs, 128, 2048), eal = make_smem_desc(act_lo + zs, 128, 2048), ebh = make_smem_desc(act_hi + os, 128, 2048);
#pragma unroll
            for (int ks = 0; ks < ST_TILE / 16; ++ks) {
                mma_bf16(d + G::N(l), eah, ebh, idesc2, ks > 0 ? 1u : acc0);
                mma_bf16(d + G::N(l), eal, ebh, idesc2, 1);                                          // the lo half of ONES is zero
                eah = desc_advance(eah, 256); eal = desc_advance(eal, 256); ebh = desc_advance(ebh, 256);
            }
        }
    } else {
        const uint32_t idesc = make_idesc_bf16_ex(128, G::K(l), 1, 1);                 // D[out, in|1] += dZ_l^T [X_l | 1 | .]
        uint64_t dah = make_smem_desc(act_hi + zs, 128, 2048), dal = make_smem_desc(act_lo + zs, 128, 2048);
        uint64_t dbh = make_smem_desc(act_hi + xs, 128, 2048), dbl = make_smem_desc(act_lo + xs, 128, 2048);
#pragma unroll
        for (int ks = 0; ks < ST_TILE / 16; ++ks) {
            mma3(d, dah, dal, dbh, dbl, idesc, ks > 0 ? 1u : acc0);
            dah = desc_advance(dah, 256); dal = desc_advance(dal, 256); dbh = desc_advance(dbh, 256); dbl = desc_advance(dbl, 256);
        }
    }
}

// write 8 fp32 values as one bf16 hi group row and one lo group row
__device__ __forceinline__ void store_split8(uint8_t* act_hi, uint8_t* act_lo, int group, int row, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pair(v[2 * q], v[2 * q + 1], h[q], l[q]);
    *reinterpret_cast<uint4*>(act_hi + group * 2048 + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(act_lo + group * 2048 + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

// forward epilogue of hidden layer l: ACC -> tanh -> X_{l+1}
template <class S, int l> __device__ __forceinline__ void epi_fwd(uint32_t tacc, uint8_t* act_hi, uint8_t* act_lo, int row, int part) {
    constexpr int og = S::out(l) / 8, gpp = (og + 3) / 4;
    static_assert(S::out(l) % 8 == 0, "hidden widths are multiples of 8");
    float v[gpp][8];
#pragma unroll
    for (int i = 0; i < gpp; ++i)
        if (part * gpp + i < og) tmem_ld_x8(tacc + (part * gpp + i) * 8, v[i]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < gpp; ++i) {
        const int g = part * gpp + i;
        if (g < og) {
#pragma unroll
            for (int k = 0; k < 8; k += 4) {                 // one reciprocal per four values: 5 MUFU instead of 8 (L1 epilogue is XU-bound)
                float y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) y[u] = v[i][k + u] * 2.8853900817779268f;
                tanh_quad_from_scaled(y, &v[i][k]);
            }
            store_split8(act_hi, act_lo, S::slot(l + 1) + g, row, v[i]);
        }
    }
}
// backward epilogue in two halves so that the wgrad MMAs (which still read X_l) overlap the arithmetic:
//   epi_bwd_compute: dX_l (ACC) * tanh'(X_l) -> registers        epi_bwd_store: registers -> dZ_{l-1}, in place over X_l
template <class S, int l> struct EpiBwd {
    static constexpr int ig = S::in(l) / 8, gpp = (ig + 3) / 4;
    static_assert(S::in(l) % 8 == 0, "hidden widths are multiples of 8");
    float v[gpp][8];
    __device__ __forceinline__ void compute(uint32_t tacc, const uint8_t* act_hi, const uint8_t* act_lo, int row, int part) {
#pragma unroll
        for (int i = 0; i < gpp; ++i)
            if (part * gpp + i < ig) tmem_ld_x8(tacc + (part * gpp + i) * 8, v[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < gpp; ++i) {
            const int g = part * gpp + i;
            if (g < ig) {
                const uint4 hh = *reinterpret_cast<const uint4*>(act_hi + (S::slot(l) + g) * 2048 + row * 16);
                const uint4 ll = *reinterpret_cast<const uint4*>(act_lo + (S::slot(l) + g) * 2048 + row * 16);
                const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float h0 = bf16lo_to_f32(hw[q]) + bf16lo_to_f32(lw[q]), h1 = bf16hi_to_f32(hw[q]) + bf16hi_to_f32(lw[q]);
                    v[i][2 * q] *= fmaf(-h0, h0, 1.f);
                    v[i][2 * q + 1] *= fmaf(-h1, h1, 1.f);
                }
            }
        }
    }
    __device__ __forceinline__ void store(uint8_t* act_hi, uint8_t* act_lo, int row, int part) {
#pragma unroll
        for (int i = 0; i < gpp; ++i)
            if (part * gpp + i < ig) store_split8(act_hi, act_lo, S::slot(l) + part * gpp + i, row, v[i]);
    }
};

template <class S> struct LayerLoop {
    // weights -> global image of the K-major hi/lo tiles: element (n, k) = W[k][n] for k < in, b[n] at k == 8 * ig (the ONES
    // feature), else 0.  e = flat element index over all layers (so every layer is built concurrently, one element per thread).
    template <int l> __device__ static void image_element(const StudentTcArgs& a, int e) {
        using G = Geo<S>;
        constexpr int K = G::K(l), N = G::N(l), in = S::in(l), out = S::out(l), kb = G::ig(l) * 8;
        if (e < N * K) {
            const int n = e % N, k = e / N;
            float v = 0.f;
            if (S::L == 4 && l == 2) {                       // SpecMLP folded layer: its non-zero elements come from fold_into_image()
                if (n < out && (k < in || k == kb)) return;
            } else if (n < out) v = k < in ? ldw(a.w[l] + k * out + n) : (k == kb ? ldw(a.b[l] + n) : 0.f);
            uint16_t h, lo;
            split_scalar(v, h, lo);
            const uint32_t off = G::wtile_off(l) + tile_off(n, k, N);
            *reinterpret_cast<uint16_t*>(a.wimg + off) = h;
            *reinterpret_cast<uint16_t*>(a.wimg + G::wtile_bytes() + off) = lo;
            return;
        }
        if constexpr (l + 1 < S::L) image_element<l + 1>(a, e - N * K);
    }
    // the image elements are dealt to the LAST threads of the grid (the fold, which is the long job, starts at the first)
    __device__ static void build_image(const StudentTcArgs& a, int gtid, int gthreads) {
        constexpr int total = Geo<S>::wtile_bytes() / 2;
        for (int e = gthreads - 1 - gtid; e < total; e += gthreads) image_element<0>(a, e);
    }
    // G_l (TMEM) -> partial gradient vector (layouts: Geo::wq)
    template <int l> __device__ static void dump_grads(const StudentTcArgs& a, uint32_t tmem, float* __restrict__ part_out, int sub, int part, int lane) {
        using G = Geo<S>;
        constexpr int in = S::in(l), out = S::out(l), kb = G::ig(l) * 8;
        const uint32_t t0 = tmem + ((uint32_t)(sub * 32) << 16) + G::gcol(l);
        if constexpr (G::wq(l)) {                                                                  // lane i = in feature (kb: bias), column j = out
            constexpr int rows = G::wq_extra(l) ? in : kb + 1;
            if (sub * 32 < rows) {                                                                 // warp-uniform
                const int i = sub * 32 + lane;
                constexpr int NC = G::N(l) / 8, PER = (NC + 3) / 4;
                float v[PER][8];
#pragma unroll
                for (int u = 0; u < PER; ++u)
                    if (part + 4 * u < NC) tmem_ld_x8(t0 + (part + 4 * u) * 8, v[u]);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int c = part + 4 * u;
                    if (c < NC) {
                        if (out % 4 == 0 && i < in && (reinterpret_cast<uintptr_t>(part_out + a.pw[l] + i * out) & 15) == 0) {   // lane's row of dW
#pragma unroll
                            for (int k = 0; k < 8; k += 4)
                                if (c * 8 + k < out)
                                    *reinterpret_cast<float4*>(part_out + a.pw[l] + i * out + c * 8 + k) = make_float4(v[u][k], v[u][k + 1], v[u][k + 2], v[u][k + 3]);
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int j = c * 8 + k;
                                if (j < out) {
                                    if (i < in) part_out[a.pw[l] + i * out + j] = v[u][k];
                                    else if (i == kb && !G::wq_extra(l)) part_out[a.pb[l] + j] = v[u][k];
                                }
                            }
                        }
                    }
                }
            }
            if constexpr (G::wq_extra(l)) {                                                        // bias gradient: lane j = out feature, first extra column
                if (part == 0 && sub * 32 < out) {
                    float v[8];
                    tmem_ld_x8(t0 + G::N(l), v);
                    tmem_ld_wait();
                    if (sub * 32 + lane < out) part_out[a.pb[l] + sub * 32 + lane] = v[0];
                }
            }
        } else {                                                                                   // lane j = out feature, column i = in feature, kb = bias
            constexpr int ncol8 = G::ig(l) + 1;                                                    // 8-column chunks incl. the bias chunk
            if (sub * 32 < out) {                                                                  // warp-uniform
                const int j = sub * 32 + lane;
                constexpr int PER = (ncol8 + 3) / 4;
                float v[PER][8];
#pragma unroll
                for (int u = 0; u < PER; ++u)
                    if (part + 4 * u < ncol8) tmem_ld_x8(t0 + (part + 4 * u) * 8, v[u]);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < PER; ++u) {
                    const int c = part + 4 * u;
                    if (c < ncol8 && j < out) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int i = c * 8 + k;
                            if (i < in) part_out[a.pw[l] + i * out + j] = v[u][k];
                            else if (i == kb) part_out[a.pb[l] + j] = v[u][k];
                        }
                    }
                }
            }
        }
        if constexpr (l + 1 < S::L) dump_grads<l + 1>(a, tmem, part_out, sub, part, lane);
    }
};

// ---- MLP folding ----------------------------------------------------------------------------------------------------------
// flat MLP parameter layout (include/reacher_b200.h): for each layer W[in][out] then b[out], layers 16-24-128-128-32-4
constexpr int M_W1 = 0, M_B1 = 384, M_W2 = 408, M_B2 = 3480, M_W3 = 3608, M_B3 = 19992, M_W4 = 20120, M_B4 = 24216, M_W5 = 24248, M_B5 = 24376,
              M_P = 24380;
// partial / reduced vector of the folded network: [dW1 db1 dW2 db2 G34(128x32) g34(32) dW5 db5 loss]
constexpr int R_W1 = 0, R_B1 = 384, R_W2 = 408, R_B2 = 3480, R_G34 = 3608, R_g34 = 7704, R_W5 = 7736, R_B5 = 7864, R_LOSS = 7868, R_N = 7872;

// SpecMLP layer 2 of the weight image: W34 = W3 W4 (128 x 32) and b34 = b3 W4 + b4, computed in fp32 (16 lanes per element, fixed
// reduction tree), split and written straight into the K-major tile (element (n = j, k = i); bias row k = 128).
__device__ __forceinline__ void fold_into_image(const float* p, uint8_t* img, int gtid, int gthreads) {
    using G = Geo<SpecMLP>;
    const int sub = gtid & 15;
    for (int idx = gtid >> 4; idx < 4128; idx += gthreads >> 4) {            // 4128 is even: both half-warps of a warp stay together
        float acc = 0.f;
        const int i = idx < 4096 ? idx >> 5 : 128, j = idx & 31;
        const float* lhs = idx < 4096 ? p + M_W3 + i * 128 : p + M_B3;
#pragma unroll
        for (int m = 0; m < 8; ++m) { const int k = sub + 16 * m; acc = fmaf(ldw(lhs + k), ldw(p + M_W4 + k * 32 + j), acc); }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0) {
            if (idx >= 4096) acc += ldw(p + M_B4 + j);
            uint16_t h, lo;
            split_scalar(acc, h, lo);
            const uint32_t off = G::wtile_off(2) + tile_off(j, i, G::N(2));
            *reinterpret_cast<uint16_t*>(img + off) = h;
            *reinterpret_cast<uint16_t*>(img + G::wtile_bytes() + off) = lo;
        }
    }
}

// reduced folded gradient -> flat MLP gradient.  dW3 = G34 W4^T and dW4 = W3^T G34 + b3 (x) g34 are dealt one row per CTA
// (row-contiguous loads); the pass-through entries and db3 go to the last threads of the grid.  `p` is a pointer with which p[M_W3 ..
// M_B4) addresses the SNAPSHOT of W3 | b3 | W4 taken when the weight image was built (k_student_image): the live parameters are being
// updated by other threads while this runs.  Every produced entry goes to sink(flat index, value) exactly once; owned_mlp() below
// enumerates the same (thread -> flat index) ownership without the arithmetic.
template <class F> __device__ __forceinline__ void finish_mlp(const float* p, const float* red, int gtid, int gthreads, F&& sink) {
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < 128; r += gridDim.x) {
        // all loads of both products are issued before the first use: one L2 round trip instead of two
        float4 g[8], w[8];
        if (tid < 128) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                g[j4] = ldw(reinterpret_cast<const float4*>(red + R_G34 + r * 32) + j4);
                w[j4] = ldw(reinterpret_cast<const float4*>(p + M_W4 + tid * 32) + j4);
            }
        }
        const int j = tid >> 4, sub = tid & 15;
        float w3[8], g4[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) { const int aa = sub + 16 * m; w3[m] = ldw(p + M_W3 + aa * 128 + r); g4[m] = ldw(red + R_G34 + aa * 32 + j); }
        const float b3r = ldw(p + M_B3 + r), g34j = ldw(red + R_g34 + j);
        float acc3 = 0.f;
        if (tid < 128) {                                             // dW3[r][k] = sum_j G34[r][j] W4[k][j],  k = tid
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                acc3 = fmaf(g[j4].x, w[j4].x, acc3); acc3 = fmaf(g[j4].y, w[j4].y, acc3); acc3 = fmaf(g[j4].z, w[j4].z, acc3); acc3 = fmaf(g[j4].w, w[j4].w, acc3);
            }
        }
        float acc4 = 0.f;                                            // dW4[r][j] = sum_a W3[a][r] G34[a][j] + b3[r] g34[j],  16 lanes per j
#pragma unroll
        for (int m = 0; m < 8; ++m) acc4 = fmaf(w3[m], g4[m], acc4);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc4 += __shfl_xor_sync(0xffffffffu, acc4, o);
        if (tid < 128) sink(M_W3 + r * 128 + tid, acc3);
        if (sub == 0) sink(M_W4 + r * 32 + j, fmaf(b3r, g34j, acc4));
    }
    constexpr int n_a = M_W3, n_b = M_W4 - M_B3, n_c = M_P + 1 - M_B4;       // [0, W3) | db3 | [db4 .. loss]
    for (int q = gthreads - 1 - gtid; q < n_a + n_b + n_c; q += gthreads) {
        if (q < n_a) sink(q, ldw(red + q));                       // dW1 db1 dW2 db2 share offsets
        else if (q < n_a + n_b) {                                    // db3[k] = sum_j W4[k][j] g34[j]
            const int k = q - n_a;
            float4 wv[8], gv[8];                                      // row k of W4 and g34 as 16 vector loads, all in flight together
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                wv[j4] = ldw(reinterpret_cast<const float4*>(p + M_W4 + k * 32) + j4);
                gv[j4] = ldw(reinterpret_cast<const float4*>(red + R_g34) + j4);
            }
            float acc = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                acc = fmaf(wv[j4].x, gv[j4].x, acc); acc = fmaf(wv[j4].y, gv[j4].y, acc); acc = fmaf(wv[j4].z, gv[j4].z, acc); acc = fmaf(wv[j4].w, gv[j4].w, acc);
            }
            sink(M_B3 + k, acc);
        } else {
            const int i = M_B4 + (q - n_a - n_b);
            sink(i, i == M_P ? ldw(red + R_LOSS) : (i < M_W5 ? ldw(red + R_g34 + (i - M_B4)) : ldw(red + R_W5 + (i - M_W5))));
        }
    }
}
// the flat indices finish_mlp() hands to its sink in THIS thread, in the same order
template <class F> __device__ __forceinline__ void owned_mlp(int gtid, int gthreads, F&& f) {
    const int tid = threadIdx.x;
    for (int r = blockIdx.x; r < 128; r += gridDim.x) {
        if (tid < 128) f(M_W3 + r * 128 + tid);
        if ((tid & 15) == 0) f(M_W4 + r * 32 + (tid >> 4));
    }
    constexpr int n_a = M_W3, n_b = M_W4 - M_B3, n_c = M_P + 1 - M_B4;
    for (int q = gthreads - 1 - gtid; q < n_a + n_b + n_c; q += gthreads) f(q < n_a ? q : (q < n_a + n_b ? M_B3 + (q - n_a) : M_B4 + (q - n_a - n_b)));
}

// phase timestamps of CTA 0 (globaltimer, ns) of the last launch -- read back with rb_debug_student_timers()
__device__ unsigned long long g_st_timers[48];
__device__ __forceinline__ void st_stamp(int i) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_st_timers[i] = t;
    }
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// env.step(s_ac) for the 512 envs of block `blk` (= 4 tiles; warp w covers 32 envs of tile 4 blk + w / 4) -- the body of k_dagger_act
// (dagger.cu; mlp_train.py:188-196).  wait: spin until the owner of that tile has published its forward pass.
__device__ __forceinline__ void act_block(const StudentTcArgs& a, int64_t blk, bool wait, uint32_t iter, int64_t ntiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = blk * 4 + (warp >> 2);
    if (tile >= ntiles) return;                                       // warp-uniform
    if (wait) {
        if (lane == 0) while (ld_acquire_gpu(a.act_flags + tile) != iter) __nanosleep(64);
        __syncwarp();
    }
    const int64_t i = blk * ST_THREADS + threadIdx.x;
    if (i >= a.B) return;
    EnvState e = load_state(a.act_qv, a.act_tp, a.act_ctr, i);
    const float4 sp = __ldcg(a.s_out + i);                            // written by another SM in this launch: L2, never L1
    bool d;
    const float r = step_env(e, sp.x, sp.y, a.act_k0, a.act_k1, a.act_offset + (uint32_t)i, d);
    store_state(a.act_qv, a.act_tp, a.act_ctr, i, e);
    a.act_prev_t[i] = __ldg(reinterpret_cast<const float4*>(a.t) + i);
    a.act_prev_rec_rew[i] = a.act_last_reward[i];
    a.act_last_reward[i] = r;
    if (a.act_rew) a.act_rew[i] = r;
    if (a.act_done) a.act_done[i] = d ? 1 : 0;
}
// end of the iteration: advance the device-side step clock, post {loss, iterations done} into host memory (one thread of the grid)
__device__ __forceinline__ void act_finish(const StudentTcArgs& a, float loss, uint32_t iter, bool bump_clock, bool post) {
    if (bump_clock) { a.act_clock[0] = iter; a.act_clock[1] += 1u; a.act_clock[2] += 1u; }
    if (post && a.act_mailbox)
        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(a.act_mailbox), "r"(__float_as_uint(loss)), "r"(iter) : "memory");
}

// The weights as the tiles want them: bf16 hi / lo split K-major tiles with the bias rows, layers 3-4 of the MLP folded (fp32), plus the
// fp32 snapshot of W3 | b3 | W4 the un-fold at the end of the step reads.  An ordinary launch in front of k_student_tc; inside
// rb_dagger_step it sits on a forked branch of the CUDA graph and runs beside the observe kernel (it only needs the parameters).
constexpr int M_SNAP = 16384 + 128 + 4096;                 // W3 | b3 | W4 are contiguous in the flat layout
template <class S> __global__ void __launch_bounds__(256) k_student_image(const StudentTcArgs a) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    if constexpr (S::L == 4) {
        fold_into_image(a.params, a.wimg, gtid, gthreads);
        for (int i = gthreads - 1 - gtid; i < M_SNAP; i += gthreads) a.snap[i] = ldw(a.params + M_W3 + i);
    }
    LayerLoop<S>::build_image(a, gtid, gthreads);
    if (gtid == gthreads - 1 && a.clock) {         // same formula as the host (student.cu: adam_lr_t), in double: ~2 us of FP64, kept off k_student_tc's path
        const double t = (double)(a.clock[1] + 1u);
        a.ctlws[0] = (float)((double)a.lr * sqrt(1.0 - pow((double)a.beta2, t)) / (1.0 - pow((double)a.beta1, t)));
    }
}

// One cooperative launch = tiles (forward, loss, backward; weight gradients in TMEM) -> partials -> grid-wide fixed-order reduction ->
// gradient of the un-folded parameters [-> peer all-reduce] [-> Adam]: two grid barriers.
template <class S>
__global__ void __launch_bounds__(ST_THREADS, 1) k_student_tc(const StudentTcArgs a) {
    using G = Geo<S>;
    constexpr int L = S::L;
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* act_hi = smem;
    uint8_t* act_lo = smem + ST_ACT_BYTES;
    uint8_t* w_hi = smem + 2 * ST_ACT_BYTES;
    uint8_t* w_lo = w_hi + G::wtile_bytes();
    StudentTcCtl& ctl = *reinterpret_cast<StudentTcCtl*>(w_lo + G::wtile_bytes());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = warp & 3, part = warp >> 2, row = sub * 32 + lane;
    const int gtid = blockIdx.x * ST_THREADS + tid, gthreads = gridDim.x * ST_THREADS;

    // ---- phase 0: TMEM, barriers, the weight image (built by k_student_image BEFORE this launch: one TMA bulk copy), zeroed activations,
    // ONES groups.  No grid-wide work and no grid barrier in front of the tiles.
    st_stamp(0);
    if (warp == 0) tmem_alloc<512>(&ctl.tmem_base);
    if (tid == 0) {
        mbar_init(&ctl.mbar, 1); mbar_init(&ctl.mbar2, 1); mbar_init(&ctl.mbar_load, 1); fence_mbar_init();
        mbar_expect_tx(&ctl.mbar_load, 2 * G::wtile_bytes());
        bulk_g2s(w_hi, a.wimg, 2 * G::wtile_bytes(), &ctl.mbar_load);
        ctl.epoch = a.epoch;
        if (a.clock) {                             // device-side step clock (CUDA-graph replay); lr_t of this step comes from k_student_image (a.ctlws)
            ctl.epoch = a.clock[2] + 1u;
            ctl.iter = a.clock[0] + 1u;
        }
    }
    for (int i = tid; i < 2 * ST_ACT_BYTES / 16; i += ST_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                               // barrier inits visible before anyone waits; zero fill complete before the ONES stores
    st_stamp(1);
    if (tid < ST_TILE) {
#pragma unroll
        for (int l = 0; l < L; ++l) *reinterpret_cast<uint16_t*>(act_hi + (S::slot(l) + G::ig(l)) * 2048 + tid * 16) = (uint16_t)0x3F80u;
    }
    mbar_wait(&ctl.mbar_load, 0);
    st_stamp(2);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = ctl.tmem_base;
    const uint32_t tacc = tmem + ((uint32_t)(sub * 32) << 16);
    const uint32_t ah = smem_u32(act_hi), al = smem_u32(act_lo), wh = smem_u32(w_hi), wl = smem_u32(w_lo);
    uint32_t phase = 0, phase2 = 0;
    float loss_acc = 0.f;
    bool first = true;

    st_stamp(3);
    const int64_t ntiles = (a.B + ST_TILE - 1) / ST_TILE;
    // software prefetch of the next tile's inputs (SpecMLP: one float4 of x -- and of x_act -- per thread) while the current tile computes
    const bool two_pass = S::IN0 == 16 && a.x_act != nullptr && !a.fwd_only;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), xav = xv;
    if constexpr (S::IN0 == 16) {
        const int64_t r0 = (int64_t)blockIdx.x * ST_TILE + (tid >> 2);
        if (blockIdx.x < ntiles && r0 < a.B) {
            xv = __ldg(reinterpret_cast<const float4*>(a.x + r0 * 16) + (tid & 3));
            if (two_pass) xav = __ldg(reinterpret_cast<const float4*>(a.x_act + r0 * 16) + (tid & 3));
        }
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * ST_TILE;
        const int nvalid = (int)min((int64_t)ST_TILE, a.B - base);
        const bool prof = tile == blockIdx.x;
#define RB_TS(i) do { if (prof) st_stamp(16 + (i)); } while (0)
        RB_TS(0);
        float4 tpd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!a.fwd_only && part == 0 && row < nvalid) tpd = __ldg(reinterpret_cast<const float4*>(a.t) + base + row);
        // pass 0 (only with x_act): acting forward on the un-dropped rows -> s_out.  pass 1: the training pass (forward, loss, backward).
#pragma unroll 1
        for (int pass = two_pass ? 0 : 1; pass < 2; ++pass) {
        const bool act_pass = pass == 0;
        const bool emit_s = two_pass ? act_pass : true;              // which forward pass produces the pdflat the caller (and the env step) sees
        // ---- X0: global fp32 -> (obfilter) -> bf16 hi/lo rows ---------------------------------------------------------------
        if constexpr (S::IN0 == 16) {
            const int r = tid >> 2, q = tid & 3;                       // 4 threads per sample row, 4 features each
            const float4 xin = act_pass ? xav : xv;
            uint32_t h0, l0, h1, l1;
            split_pair(xin.x, xin.y, h0, l0);
            split_pair(xin.z, xin.w, h1, l1);
            const uint32_t off = (S::slot(0) + (q >> 1)) * 2048 + r * 16 + (q & 1) * 8;
            *reinterpret_cast<uint2*>(act_hi + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(act_lo + off) = make_uint2(l0, l1);
            if (!act_pass) {
                const int64_t rn = (tile + gridDim.x) * ST_TILE + r;
                xv = make_float4(0.f, 0.f, 0.f, 0.f); xav = xv;
                if (rn < a.B) {
                    xv = __ldg(reinterpret_cast<const float4*>(a.x + rn * 16) + q);
                    if (two_pass) xav = __ldg(reinterpret_cast<const float4*>(a.x_act + rn * 16) + q);
                }
            }
        } else {
            for (int e = tid; e < ST_TILE * S::IN0; e += ST_THREADS) {
                const int r = e / S::IN0, f = e - r * S::IN0;
                float v = 0.f;
                if (r < nvalid) {
                    v = __ldg(a.x + base * S::IN0 + e);
                    if (S::OBFILTER) v = fminf(5.f, fmaxf(-5.f, (v - __ldg(a.obf + f)) / __ldg(a.obf + 11 + f)));
                }
                uint16_t h, lo;
                split_scalar(v, h, lo);
                const uint32_t off = (S::slot(0) + (f >> 3)) * 2048 + r * 16 + (f & 7) * 2;
                *reinterpret_cast<uint16_t*>(act_hi + off) = h;
                *reinterpret_cast<uint16_t*>(act_lo + off) = lo;
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        RB_TS(1);

        // ---- forward ----------------------------------------------------------------------------------------------------
#define RB_ST_FWD(l)                                                               \
        if constexpr (l < L) {                                                     \
            if (warp == 0 && elect_one_sync()) {                                   \
                fence_after_sync();                                                \
                issue_fwd<S, (l < L ? l : 0)>(tmem, ah, al, wh, wl);               \
                mma_commit(&ctl.mbar);                                             \
            }                                                                      \
            mbar_wait(&ctl.mbar, phase); phase ^= 1u;                              \
            fence_after_sync();                                                    \
            RB_TS(2 + 2 * l);                                                      \
            if constexpr (l < L - 1) {                                             \
                epi_fwd<S, (l < L - 1 ? l : 0)>(tacc, act_hi, act_lo, row, part);  \
                fence_async_smem();                                                \
                fence_before_sync();                                               \
                __syncthreads();                                                   \
                RB_TS(3 + 2 * l);                                                  \
            }                                                                      \
        }
        RB_ST_FWD(0) RB_ST_FWD(1) RB_ST_FWD(2) RB_ST_FWD(3)
#undef RB_ST_FWD
        // ---- output epilogue: s = ACC[:, 0:4]; loss and dL/ds (loss.py:8-13; reverse: backup/student_rollout.py:639-640; squared errors) ---
        if (part == 0) {
            float s[4];
            tmem_ld_x4(tacc, s);
            tmem_ld_wait();
            if (emit_s && row < nvalid && a.s_out) a.s_out[base + row] = make_float4(s[0], s[1], s[2], s[3]);
            if (!a.fwd_only && !act_pass) {
                float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (row < nvalid) {
                    float4 dd;
                    loss_acc += pd_loss_row(make_float4(s[0], s[1], s[2], s[3]), tpd, a.loss_kind, dd);
                    d[0] = dd.x; d[1] = dd.y; d[2] = dd.z; d[3] = dd.w;
                }
                store_split8(act_hi, act_lo, 0, row, d);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        RB_TS(10);
        // s_out of this tile is published (release: cumulative over the CTA barrier above).  Warp 1 does it: warp 0 issues the next MMAs,
        // and the other warps only wait for them, so the fence latency is hidden.
        if (a.act_on && emit_s && tid == 32) st_release_gpu(a.act_flags + tile, ctl.iter);
        }   // pass
        if (a.fwd_only) continue;

        // ---- backward: wgrad_l (accumulates in TMEM) + dgrad_l, then dZ_{l-1} in place over X_l ------------------------------
#define RB_ST_BWD(l)                                                               \
        if constexpr (l < L) {                                                     \
            if (warp == 0 && elect_one_sync()) {                                   \
                fence_after_sync();                                                \
                if constexpr (l > 0) {                                             \
                    issue_dgrad<S, (l < L ? l : 0)>(tmem, ah, al, wh, wl);         \
                    mma_commit(&ctl.mbar);                                         \
                }                                                                  \
                issue_wgrad<S, (l < L ? l : 0)>(tmem, ah, al, first ? 1u : 0u);    \
                mma_commit(&ctl.mbar2);                                            \
            }                                                                      \
            if constexpr (l > 0) {                                                 \
                EpiBwd<S, (l > 0 && l < L ? l : 1)> eb;                            \
                mbar_wait(&ctl.mbar, phase); phase ^= 1u;                          \
                fence_after_sync();                                                \
                eb.compute(tacc, act_hi, act_lo, row, part);                       \
                mbar_wait(&ctl.mbar2, phase2); phase2 ^= 1u;                       \
                eb.store(act_hi, act_lo, row, part);                               \
                fence_async_smem();                                                \
            } else {                                                               \
                mbar_wait(&ctl.mbar2, phase2); phase2 ^= 1u;                       \
            }                                                                      \
            fence_before_sync();                                                   \
            __syncthreads();                                                       \
            RB_TS(14 - l);                                                         \
        }
        RB_ST_BWD(3) RB_ST_BWD(2) RB_ST_BWD(1) RB_ST_BWD(0)
#undef RB_ST_BWD
        first = false;
    }

    st_stamp(4);
    if (!a.fwd_only) {
        // ---- this CTA's partial gradient + loss -> global ---------------------------------------------------------------------
        float* part_out = a.partials + (size_t)blockIdx.x * a.pstride;
        fence_after_sync();
        if (!first) LayerLoop<S>::template dump_grads<0>(a, tmem, part_out, sub, part, lane);
        float v = loss_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) ctl.red[warp] = v;
        __syncthreads();
        if (tid == 0) part_out[a.ploss] = (ctl.red[0] + ctl.red[1]) + (ctl.red[2] + ctl.red[3]);     // part 0 = warps 0..3
        // ---- fused env step, part 1: CTAs [r, grid) own one tile less than the others (r = ntiles mod grid) and would idle for one tile
        // period (~10 us ~ two 512-env blocks) at the barrier below; they step blocks e, e + n_early of the env range instead.
        const int64_t act_nblocks = (a.B + ST_THREADS - 1) / ST_THREADS;
        int64_t act_covered = 0;
        if (a.act_on) {
            const int r = (int)(ntiles % gridDim.x), n_early = r ? (int)gridDim.x - r : 0;
            act_covered = min(act_nblocks, (int64_t)2 * n_early);
            if (r && (int)blockIdx.x >= r)
                for (int64_t blk = (int)blockIdx.x - r; blk < act_covered; blk += n_early) act_block(a, blk, true, ctl.iter, ntiles);
        }
        st_stamp(5);
        grid.sync();
        st_stamp(6);
        // (envs beyond act_covered blocks are stepped by a k_dagger_act launch behind this kernel: 64 warps per SM hide the latency of the
        // physics chain better than the 16 of this kernel)
        // ---- grid-wide reduction of the partial vectors in CTA order with a fixed tree => bit-reproducible.  A warp owns 32
        // consecutive elements (coalesced 128-byte reads of every partial); 8 warps split the partials, shared memory combines them.
        {
            float* red2 = reinterpret_cast<float*>(act_hi);                 // 16 x 32 floats of scratch (activations are dead here)
            const int nparts = (int)min((int64_t)gridDim.x, ntiles), sl = warp & 7;
            const int nblk = (a.pstride + 31) / 32;
            for (int eb0 = blockIdx.x * 2; eb0 < nblk; eb0 += gridDim.x * 2) {
                const int i = (eb0 + (warp >> 3)) * 32 + lane;
                float acc = 0.f;
                if (i < a.pstride) {
                    float pv[ST_MAX_GRID / 8];                         // all partials of this slot in flight together (one L2 round trip)
#pragma unroll
                    for (int u = 0; u < ST_MAX_GRID / 8; ++u) {
                        const int b = sl + 8 * u;
                        pv[u] = b < nparts ? ldw(a.partials + (size_t)b * a.pstride + i) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < ST_MAX_GRID / 8; ++u) acc += pv[u];      // CTA order, absent partials add 0
                }
                red2[warp * 32 + lane] = acc;
                __syncthreads();
                if (sl == 0 && i < a.pstride) {
                    const float* r = red2 + (warp >> 3) * 256 + lane;
                    a.red[i] = ((r[0] + r[32]) + (r[64] + r[96])) + ((r[128] + r[160]) + (r[192] + r[224]));
                }
                __syncthreads();
            }
        }
        st_stamp(7);
        grid.sync();
        st_stamp(8);
        // ---- last phase, NO further grid barrier: gradient of the un-folded parameters -> [data parallel: one-shot all-reduce over NVLink peer
        // memory] -> TF-form Adam, element by element.  Whoever produces gradient entry i also exchanges it and updates parameter i, and the
        // un-fold reads the W3 | b3 | W4 snapshot (a.snap) instead of the live parameters, so no thread waits for another one of this GPU.
        const uint32_t epoch = ctl.epoch;
        float* gl = a.gradloss;
        const float lr_t = a.clock ? ldw(a.ctlws) : a.lr_t;
        auto apply = [&](int i, float g) {                   // final value of entry i: gradient vector, loss mailbox, Adam (same arithmetic as k_adam)
            gl[i] = g;
            if (i == a.P) { if (a.act_on) act_finish(a, g, ctl.iter, a.world <= 1, true); }
            else if (a.do_adam) {
                float pi = a.adam_p[i], mi = a.adam_m[i], vi = a.adam_v[i];
                adam_update(pi, mi, vi, g, lr_t, a.beta1, a.beta2, a.eps, a.gscale);
                a.adam_p[i] = pi; a.adam_m[i] = mi; a.adam_v[i] = vi;
            }
        };
        if (a.world <= 1) {
            if constexpr (S::L == 4) finish_mlp(a.snap - M_W3, a.red, gtid, gthreads, apply);
            else
                for (int i = gtid; i <= a.P; i += gthreads) apply(i, ldw(a.red + i));
            st_stamp(9);
        } else {
            // One-shot all-reduce (MpiAdam.update's Allreduce, backup/student_rollout.py:709), low-latency form: every rank PUSHES {value, epoch}
            // pairs (one 8-byte store each, delivered atomically) into every rank's receive area, then sums its own area in rank order as the
            // pairs of this epoch arrive -- no flag round trip, no remote loads; the one-way NVLink latency overlaps the pushes still in flight.
            // Areas are double buffered by epoch parity (a fast rank may push step k + 1 while a slow one still reads step k).
            const int SL = (a.P + 1 + 63) / 64 * 64;                               // slot stride (elements) inside a receive area
            uint2* const* ll = a.clock ? a.peer_ll2[epoch & 1u] : a.peer_ll;
            auto push = [&](int i, float g) {
                const uint32_t v = __float_as_uint(g);
                for (int q = 0; q < a.world; ++q) st_ll(ll[q] + (size_t)a.rank * SL + i, v, epoch);
            };
            if constexpr (S::L == 4) finish_mlp(a.snap - M_W3, a.red, gtid, gthreads, push);
            else
                for (int i = gtid; i <= a.P; i += gthreads) push(i, ldw(a.red + i));
            st_stamp(34);
            const uint2* mine = ll[a.rank];
            // All `world` pairs of an element are requested TOGETHER (one L2 round trip instead of `world` serial ones: polling rank by rank cost
            // 8 x 0.7 us at 8 ranks), re-read only while a pair still carries an older epoch, then added in rank order: bit-identical sums on every rank.
            auto recv = [&](int i) {
                uint2 w[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) w[r] = r < a.world ? ld_ll(mine + (size_t)r * SL + i) : make_uint2(0u, epoch);
                for (;;) {
                    bool ready = true;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (w[r].y != epoch) { w[r] = ld_ll(mine + (size_t)r * SL + i); ready = ready && w[r].y == epoch; }
                    }
                    if (ready) break;
                }
                float tot = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) if (r < a.world) tot += __uint_as_float(w[r].x);
                apply(i, tot);
            };
            if constexpr (S::L == 4) owned_mlp(gtid, gthreads, recv);
            else
                for (int i = gtid; i <= a.P; i += gthreads) recv(i);
            st_stamp(9);
            if (a.act_on && gtid == 0) act_finish(a, 0.f, ctl.iter, true, false);     // every CTA read the clock before the first barrier
        }
    }
    st_stamp(10);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
    st_stamp(11);
}

template <class S> static size_t student_tc_smem() { return 2 * (size_t)ST_ACT_BYTES + 2 * (size_t)Geo<S>::wtile_bytes() + sizeof(StudentTcCtl); }

static int tc_grid(int64_t B, int* grid) {
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    // always one CTA per SM: CTAs without a tile still share the grid-wide phases (fold, weight image, reduction, un-fold, Adam),
    // which would otherwise run on a handful of CTAs for small batches (a 128-sample step took 630 us on a single CTA)
    (void)B;
    *grid = min(sms, ST_MAX_GRID);
    return RB_OK;
}

// workspace (floats): [snapshot W3 | b3 | W4 20608][hand-over scalars 64][reduced 8192][weight image 12288 (48 KB)][partials ST_MAX_GRID * 8192]
constexpr size_t WS_SNAP = 0, WS_CTL = 20608, WS_RED = WS_CTL + 64, WS_IMG = WS_RED + 8192, WS_PART = WS_IMG + 12288, WS_PSTRIDE_MAX = 8192;
static_assert(M_SNAP <= (int)WS_CTL && (WS_IMG * 4) % 16 == 0 && M_W3 + 16384 == M_B3 && M_B3 + 128 == M_W4 && M_W4 + 4096 == M_B4, "W3 | b3 | W4 snapshot layout");

size_t student_tc_workspace_floats() { return WS_PART + (size_t)ST_MAX_GRID * WS_PSTRIDE_MAX; }

template <class S> static int launch_student_image(const StudentTcArgs& a, cudaStream_t st) {
    k_student_image<S><<<132, 256, 0, st>>>(a);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}
template <class S> static int launch_student_tc(StudentTcArgs& a, int grid, cudaStream_t st, bool image_prebuilt) {
    static_assert(2 * Geo<S>::wtile_bytes() <= 12288 * 4, "weight image does not fit its workspace slot");
    static_assert((2 * Geo<S>::wtile_bytes()) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
    if (!image_prebuilt) { const int rc = launch_student_image<S>(a, st); if (rc) return rc; }
    const size_t smem = student_tc_smem<S>();
    static bool seen[RB_MAX_DEVICES] = {};             // function attributes and occupancy are per device
    if (first_use_on_device(seen)) {
        RB_CUDA(cudaFuncSetAttribute(k_student_tc<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;                                   // a cooperative grid must be co-resident: one CTA per SM has to fit
        RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_student_tc<S>, ST_THREADS, smem));
        RB_REQUIRE(occ >= 1, "k_student_tc does not fit on this device (shared memory / registers)");
    }
    void* args[] = {(void*)&a};
    RB_CUDA(cudaLaunchCooperativeKernel((const void*)k_student_tc<S>, dim3(grid), dim3(ST_THREADS), args, smem, st));
    return RB_OK;
}

struct AdamFuse { float* p; float* m; float* v; float lr_t, beta1, beta2, eps, gscale; };
struct PeerExchange { int world, rank; uint32_t epoch; const uint64_t* gl_ptrs; const uint64_t* flag_ptrs; const uint64_t* gl_ptrs_alt; };
struct StepClock { const uint32_t* clock; float lr, beta1, beta2; };
struct ActFuse {              // fused env step of rb_dagger_step (see StudentTcArgs::act_*); x_act: un-dropped input rows or NULL
    float4* qv; float4* tp; uint4* ctr; float4* prev_t; float* prev_rec_rew; float* last_reward; float* rew; uint8_t* done;
    uint32_t k0, k1, offset; uint32_t* flags; uint32_t* clock; uint2* mailbox; const float* x_act;
    int image_prebuilt;       // the caller already ran student_tc_build_image() on these parameters (rb_dagger_step: on a forked graph branch)
};

int student_tc_run_ex(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only, float* s_out,
                      float* gradloss, void* workspace, const AdamFuse* adam, const PeerExchange* px, const StepClock* clk, const ActFuse* act,
                      cudaStream_t st);
// number of leading envs the fused env step of a B-sample launch covers (same arithmetic as the kernel: two blocks per early CTA)
int64_t student_tc_act_covered(int64_t B, int grid) {
    const int64_t ntiles = (B + ST_TILE - 1) / ST_TILE, nblocks = (B + ST_THREADS - 1) / ST_THREADS;
    const int r = (int)(ntiles % grid), n_early = r ? grid - r : 0;
    const int64_t covered = nblocks < (int64_t)2 * n_early ? nblocks : (int64_t)2 * n_early;
    const int64_t envs = covered * ST_THREADS;
    return envs < B ? envs : B;
}
int student_tc_grid(int* grid) { return tc_grid(0, grid); }
int student_tc_run(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only, float* s_out,
                   float* gradloss, void* workspace, const AdamFuse* adam, const PeerExchange* px, const StepClock* clk, cudaStream_t st) {
    return student_tc_run_ex(kind, params, x, tpd, B, loss_kind, fwd_only, s_out, gradloss, workspace, adam, px, clk, nullptr, st);
}

// weight image (+ un-fold snapshot) of `params` into the workspace: what every k_student_tc launch on these parameters starts from
int student_tc_build_image(int kind, const float* params, void* workspace, const StepClock* clk, cudaStream_t st) {
    RB_REQUIRE(params != nullptr && workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "bad arguments");
    float* ws = (float*)workspace;
    StudentTcArgs a{};
    a.params = params; a.snap = ws + WS_SNAP; a.ctlws = ws + WS_CTL; a.wimg = (uint8_t*)(ws + WS_IMG);
    if (clk && clk->clock) { a.clock = clk->clock; a.lr = clk->lr; a.beta1 = clk->beta1; a.beta2 = clk->beta2; }
    if (kind == RB_STUDENT_MLP) {
        a.w[0] = params + M_W1; a.b[0] = params + M_B1; a.w[1] = params + M_W2; a.b[1] = params + M_B2;
        a.w[2] = nullptr; a.b[2] = nullptr; a.w[3] = params + M_W5; a.b[3] = params + M_B5;
        return launch_student_image<SpecMLP>(a, st);
    }
    const PolicyOffsets o = policy_offsets(4);
    a.w[0] = params + o.W1; a.b[0] = params + o.b1; a.w[1] = params + o.W2; a.b[1] = params + o.b2; a.w[2] = params + o.W3; a.b[2] = params + o.b3;
    return launch_student_image<SpecPOL>(a, st);
}

int student_tc_run_ex(int kind, const float* params, const float* x, const float* tpd, int64_t B, int loss_kind, int fwd_only, float* s_out,
                      float* gradloss, void* workspace, const AdamFuse* adam, const PeerExchange* px, const StepClock* clk, const ActFuse* act,
                      cudaStream_t st) {
    RB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(params) & 3) == 0, "x must be 16-byte aligned");
    RB_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    float* ws = (float*)workspace;
    int grid = 1;
    int rc = tc_grid(B, &grid);
    if (rc) return rc;
    StudentTcArgs a{};
    a.x = x; a.t = tpd; a.s_out = (float4*)s_out; a.B = B; a.loss_kind = loss_kind; a.fwd_only = fwd_only; a.partials = ws + WS_PART;
    a.params = params; a.snap = ws + WS_SNAP; a.ctlws = ws + WS_CTL; a.wimg = (uint8_t*)(ws + WS_IMG); a.red = ws + WS_RED; a.gradloss = gradloss;
    if (adam && !fwd_only) {
        a.do_adam = 1; a.adam_p = adam->p; a.adam_m = adam->m; a.adam_v = adam->v;
        a.lr_t = adam->lr_t; a.beta1 = adam->beta1; a.beta2 = adam->beta2; a.eps = adam->eps; a.gscale = adam->gscale;
    }
    a.world = 1;
    if (px && !fwd_only && px->world > 1) {
        RB_REQUIRE(px->world <= 8 && px->rank >= 0 && px->rank < px->world, "peer exchange supports 2..8 ranks");
        a.world = px->world; a.rank = px->rank; a.epoch = px->epoch;
        for (int r = 0; r < px->world; ++r) {
            a.peer_ll[r] = (uint2*)px->gl_ptrs[r];
            a.peer_ll2[0][r] = (uint2*)px->gl_ptrs[r];                                  // receive areas of even epochs
            a.peer_ll2[1][r] = (uint2*)(px->gl_ptrs_alt ? px->gl_ptrs_alt[r] : px->gl_ptrs[r]);   // ... of odd epochs
        }
    }
    if (clk && clk->clock) { a.clock = clk->clock; a.lr = clk->lr; }
    if (act) {
        RB_REQUIRE(!fwd_only && adam && a.clock && s_out && tpd && act->flags && act->clock == a.clock, "fused env step needs the full optimiser step with the device clock");
        a.act_on = 1; a.act_qv = act->qv; a.act_tp = act->tp; a.act_ctr = act->ctr; a.act_prev_t = act->prev_t; a.act_prev_rec_rew = act->prev_rec_rew;
        a.act_last_reward = act->last_reward; a.act_rew = act->rew; a.act_done = act->done; a.act_k0 = act->k0; a.act_k1 = act->k1; a.act_offset = act->offset;
        a.act_flags = act->flags; a.act_clock = act->clock; a.act_mailbox = act->mailbox;
        if (kind == RB_STUDENT_MLP) a.x_act = act->x_act;
    }
    if (kind == RB_STUDENT_MLP) {
        a.w[0] = params + M_W1; a.b[0] = params + M_B1; a.w[1] = params + M_W2; a.b[1] = params + M_B2;
        a.w[2] = nullptr; a.b[2] = nullptr; a.w[3] = params + M_W5; a.b[3] = params + M_B5;
        a.pw[0] = R_W1; a.pb[0] = R_B1; a.pw[1] = R_W2; a.pb[1] = R_B2; a.pw[2] = R_G34; a.pb[2] = R_g34; a.pw[3] = R_W5; a.pb[3] = R_B5;
        a.ploss = R_LOSS; a.pstride = R_N; a.P = M_P;            // R_N: multiple of 4 floats, keeps every CTA's partial 16-byte aligned
        return launch_student_tc<SpecMLP>(a, grid, st, act && act->image_prebuilt);
    }
    const PolicyOffsets o = policy_offsets(4);
    a.obf = params;
    a.w[0] = params + o.W1; a.b[0] = params + o.b1; a.w[1] = params + o.W2; a.b[1] = params + o.b2; a.w[2] = params + o.W3; a.b[2] = params + o.b3;
    a.pw[0] = o.W1; a.pb[0] = o.b1; a.pw[1] = o.W2; a.pb[1] = o.b2; a.pw[2] = o.W3; a.pb[2] = o.b3;
    a.ploss = o.total; a.pstride = o.total + 1; a.P = o.total;
    // obfilter / logstd entries of the partial vectors are never written by the kernel: keep them zero
    if (!fwd_only) RB_CUDA(cudaMemsetAsync(ws + WS_PART, 0, sizeof(float) * (size_t)grid * a.pstride, st));
    return launch_student_tc<SpecPOL>(a, grid, st, act && act->image_prebuilt);
}

}  // namespace rb

// debug: phase timestamps (ns) of CTA 0 of the last k_student_tc launch: 0 start, 1 image built, 2 after sync, 3 image loaded, 4 tiles done,
// 5 partials written, 6 after sync, 7 reduced, 8 after sync, 9 un-folded, 10 Adam done, 11 end
extern "C" int rb_debug_student_timers(unsigned long long* host_out16) {   // 48 values: [0,12) launch phases, [16,31) first tile of CTA 0
    RB_REQUIRE(host_out16 != nullptr, "NULL argument");
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpyFromSymbol(host_out16, rb::g_st_timers, sizeof(unsigned long long) * 48));
    return RB_OK;
}

extern "C" int rb_student_mode_available(int mode) { return mode == RB_MODE_FP32 || mode == RB_MODE_TC; }
extern "C" int rb_mode_available(int mode) { return mode == RB_MODE_FP32 || mode == RB_MODE_TC; }
