// RB_MODE_TC: the baselines MlpPolicy (11 -> 64 tanh -> 64 tanh -> nout) on the 5th-gen tensor cores.
// Replaces sess.run((pi.pd.mean, pi.pd.flat)) (/root/reference src/distilation/mlp_train.py:123-125,165-167; network
// teacher.py:12-16) and, fused with the env step, the teacher warm-up loop mlp_train.py:120-139.
//
// Work decomposition.  A *tile* = 128 envs/samples = the 128 TMEM lanes of one M=128 accumulator = 4 warps; thread r of a
// tile owns row r end to end (env state in its registers, accumulator row through tcgen05.ld 32x32b).  A CTA holds NT tiles
// that share ONE shared-memory copy of the split weights; each tile has its own A-operand buffers, its own TMEM
// columns, its own mbarrier and its own named barrier, so tiles run unsynchronised and hide each other's MMA / MUFU
// latency.  One CTA per SM; envs are dealt to CTAs in units of one warp (32 envs) so every SM gets the same number of
// warps +-1 whatever N is (65 536 envs = 2048 warp-units = 13.8 per SM; a 128-env granularity would leave a 15 % tail).
//
// Per policy evaluation a tile runs two dependent GEMMs [128 x K] * [K x N] (A, B in shared memory, D in TMEM) + a fused output layer:
//     L1: K = 16 (11 obs + a ones column that carries b1 + zero pad), N = 64
//     L2: K = 64 (+ one extra K = 16 step: a constant ones-tile times the b2 row), N = 64
//     L3: 64 -> nout on the CUDA cores, fused into the L2 epilogue (fp32 FMAs; no third GEMM round trip)
// Each tile owns two 64-column TMEM accumulators (when 128 columns per tile fit, i.e. up to 4 tiles per CTA): the L2 MMAs of K-steps
// 0-1 are issued as soon as the first half of the L1 epilogue has written them, so most of the L2 GEMM runs under the second half of
// the L1 epilogue instead of after it (the MMA round trip sits in the serial chain of every env step).
// Operands are bf16 hi/lo splits of the fp32 values ("bf16x3": A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, fp32 accumulate in
// TMEM), which keeps the result within ~1e-5 of the fp32 network.  W1/b1/W2/b2 are pre-multiplied by 2*log2(e) when they
// are split, so the hidden epilogue is  tanh = 1 - 2 / (ex2(acc) + 1)  with ONE reciprocal per four elements (1.25 MUFU per
// element: the XU pipe sits inside the serial chain of every env step) and packed FADD2 / FMUL2 / FFMA2 arithmetic, then the
// bf16 hi/lo split written straight into the next layer's A tile (no-swizzle K-major layout, see tc_common.cuh).
#include <cstddef>

#include "common.cuh"
#include "dagger_input.cuh"
#include "physics.cuh"
#include "tc_common.cuh"

#ifndef RB_TANH_VARIANT
#define RB_TANH_VARIANT 5       // 5: four values per reciprocal, packed FADD2 / FMUL2 / FFMA2 (default); 4: pair per reciprocal, packed;
                                // 1: pair, scalar ops; 0 / 2 / 3: experiments
#endif

namespace rb {

using namespace tc;

constexpr int TILE = 128;               // rows per tile == threads per tile
constexpr int MAX_TILES = 6;
constexpr float TANH_PRESCALE = 2.8853900817779268f;   // 2 * log2(e)

struct __align__(128) TcShared {        // one per CTA: split weights + per-tile barriers
    uint8_t B2_hi[64 * 64 * 2];         // [8 chunks][64 rows (n)][16 B]
    uint8_t B2_lo[64 * 64 * 2];
    uint8_t B2b_hi[64 * 16 * 2];        // bias K-step of layer 2: k = 0 carries b2, k = 1..15 zero
    uint8_t B2b_lo[64 * 16 * 2];
    uint8_t B1_hi[64 * 16 * 2];
    uint8_t B1_lo[64 * 16 * 2];
    uint8_t ONES[128 * 16 * 2];         // A operand of the bias K-step: column 0 = 1.0, rest 0
    float W3f[64 * 4];                  // output layer in fp32, [k][nout]: applied on the CUDA cores
    float b3[4];
    float mu[12];
    float inv_sd[12];
    float logstd[4];
    uint64_t mbar[MAX_TILES];           // ---- everything above this line is the "weight image" (built once, bulk-copied by TMA)
    uint64_t mbar_load;
    uint32_t tmem_base;
};
constexpr uint32_t TC_IMAGE_BYTES = (uint32_t)offsetof(TcShared, mbar);
static_assert(TC_IMAGE_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
struct __align__(128) TcTile {
    uint8_t A_hi[128 * 64 * 2];         // [8 chunks][128 rows][16 B]; the L1 tile (2 chunks) aliases its head; bytes 8192.. double as obs strips
    uint8_t A_lo[128 * 64 * 2];
};
constexpr uint32_t STRIP_OFF = 8192;    // obs staging strips live in A_hi[8192 .. 8192 + 4*1408)

// TMEM columns per tile: 128 (layer-1 accumulator in [0,64), layer-2 accumulator in [64,128): lets the layer-2 MMAs start while the layer-1
// epilogue is still reading) when NT tiles fit the 512 columns, else one shared 64-column accumulator
template <int NT> __host__ __device__ constexpr int tile_cols() { return NT * 128 <= 512 ? 128 : 64; }
template <int NT> __host__ __device__ constexpr int tmem_cols() {
    return NT * tile_cols<NT>() <= 64 ? 64 : (NT * tile_cols<NT>() <= 128 ? 128 : (NT * tile_cols<NT>() <= 256 ? 256 : 512));
}
template <int NT> constexpr size_t tc_smem_bytes() { return sizeof(TcShared) + (size_t)NT * sizeof(TcTile); }

__device__ __forceinline__ void tile_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// split W (fp32, row-major [K][N] in global) into the K-major B tiles (element (n,k) = W[k][n]); all threads of the CTA
__device__ inline void policy_tc_load_weights(TcShared& S, const float* __restrict__ p, int nout) {
    const PolicyOffsets o = policy_offsets(nout);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 64 * 16; i += nt) {          // B1: n = i % 64, k = i / 64 ; k == 11 carries the bias
        const int n = i & 63, k = i >> 6;
        const float w = k < 11 ? __ldg(p + o.W1 + k * 64 + n) : (k == 11 ? __ldg(p + o.b1 + n) : 0.f);
        uint16_t h, l;
        split_scalar(w * TANH_PRESCALE, h, l);
        *reinterpret_cast<uint16_t*>(S.B1_hi + tile_off(n, k, 64)) = h;
        *reinterpret_cast<uint16_t*>(S.B1_lo + tile_off(n, k, 64)) = l;
        const float bb = k == 0 ? __ldg(p + o.b2 + n) * TANH_PRESCALE : 0.f;   // B2b: k == 0 carries b2
        split_scalar(bb, h, l);
        *reinterpret_cast<uint16_t*>(S.B2b_hi + tile_off(n, k, 64)) = h;
        *reinterpret_cast<uint16_t*>(S.B2b_lo + tile_off(n, k, 64)) = l;
    }
    for (int i = tid; i < 64 * 64; i += nt) {
        const int n = i & 63, k = i >> 6;
        uint16_t h, l;
        split_scalar(__ldg(p + o.W2 + k * 64 + n) * TANH_PRESCALE, h, l);
        *reinterpret_cast<uint16_t*>(S.B2_hi + tile_off(n, k, 64)) = h;
        *reinterpret_cast<uint16_t*>(S.B2_lo + tile_off(n, k, 64)) = l;
    }
    for (int i = tid; i < 64 * 4; i += nt) S.W3f[i] = i < 64 * nout ? __ldg(p + o.W3 + i) : 0.f;     // [k][nout], as in the parameter vector
    for (int i = tid; i < 128 * 16; i += nt) {
        const int r = i & 127, k = i >> 7;
        *reinterpret_cast<uint16_t*>(S.ONES + tile_off(r, k, 128)) = k == 0 ? (uint16_t)0x3F80u : (uint16_t)0u;
    }
    if (tid < 4) {
        S.b3[tid] = tid < nout ? __ldg(p + o.b3 + tid) : 0.f;
        S.logstd[tid] = tid < 2 ? __ldg(p + o.logstd + tid) : 0.f;
    }
    if (tid < 12) {
        S.mu[tid] = tid < 11 ? __ldg(p + o.mu + tid) : 0.f;
        S.inv_sd[tid] = tid < 11 ? 1.0f / __ldg(p + o.sd + tid) : 0.f;
    }
}

// D[tmem_d] (+)= sum over k-steps and hi/lo terms of A * B^T.  Called by ONE thread.  A tiles have 128 rows, B tiles n_rows_b.
__device__ __forceinline__ void issue_bf16x3(uint32_t tmem_d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int ksteps, int n_rows_b,
                                             uint32_t idesc, uint32_t acc, int ks0 = 0) {
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)n_rows_b * 16;
    uint64_t dah = make_smem_desc(ah + ks0 * 2 * a_lbo, a_lbo, 128), dal = make_smem_desc(al + ks0 * 2 * a_lbo, a_lbo, 128);
    uint64_t dbh = make_smem_desc(bh + ks0 * 2 * b_lbo, b_lbo, 128), dbl = make_smem_desc(bl + ks0 * 2 * b_lbo, b_lbo, 128);
#pragma unroll
    for (int ks = 0; ks < ksteps; ++ks) {
        mma_bf16(tmem_d, dah, dbh, idesc, acc);
        mma_bf16(tmem_d, dal, dbh, idesc, 1);
        mma_bf16(tmem_d, dah, dbl, idesc, 1);
        acc = 1;
        dah = desc_advance(dah, 2 * a_lbo); dal = desc_advance(dal, 2 * a_lbo); dbh = desc_advance(dbh, 2 * b_lbo); dbl = desc_advance(dbl, 2 * b_lbo);
    }
}

// 16 accumulator columns (already scaled by 2 log2 e) -> tanh -> bf16 hi/lo -> chunks 2*cc, 2*cc+1 of this thread's A row
__device__ __forceinline__ void hidden_chunk(TcTile& T, const float* v, int cc, int row) {
#pragma unroll
    for (int q8 = 0; q8 < 2; ++q8) {
        uint32_t h[4], l[4];
#if RB_TANH_VARIANT == 5
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            f32x2 t10, t32;
            tanh_quad_packed(v + 8 * q8 + 4 * q, t10, t32);
            split_packed_swapped(t10, h[2 * q], l[2 * q]);
            split_packed_swapped(t32, h[2 * q + 1], l[2 * q + 1]);
        }
#elif RB_TANH_VARIANT == 3
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float t[4];
            tanh_quad_from_scaled(v + 8 * q8 + 4 * q, t);
            split_pair(t[0], t[1], h[2 * q], l[2 * q]);
            split_pair(t[2], t[3], h[2 * q + 1], l[2 * q + 1]);
        }
#else
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#if RB_TANH_VARIANT == 4
            tanh_split_pair_packed(v[8 * q8 + 2 * q], v[8 * q8 + 2 * q + 1], h[q], l[q]);
#else
            float t0, t1;
#if RB_TANH_VARIANT == 1
            tanh_pair_from_scaled(v[8 * q8 + 2 * q], v[8 * q8 + 2 * q + 1], t0, t1);
#elif RB_TANH_VARIANT == 2
            t0 = tanh_from_scaled_newton(v[8 * q8 + 2 * q]); t1 = tanh_from_scaled_newton(v[8 * q8 + 2 * q + 1]);
#else
            t0 = tanh_from_scaled(v[8 * q8 + 2 * q]); t1 = tanh_from_scaled(v[8 * q8 + 2 * q + 1]);
#endif
            split_pair(t0, t1, h[q], l[q]);
#endif
        }
#endif
        *reinterpret_cast<uint4*>(T.A_hi + (2 * cc + q8) * 2048 + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(T.A_lo + (2 * cc + q8) * 2048 + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// W3f layout: [k][NOUT] (nout == 2: one 128-bit load brings the weight pairs of k and k + 1)
template <int NOUT> __device__ __forceinline__ void final_pair(const TcShared& S, int k, float t0, float t1, f32x2* o) {
    // (o0, o1) += t (w[k][0], w[k][1]): one FFMA2 per k (and one more for (o2, o3) when NOUT == 4); same order of additions as a scalar loop over k
    if (NOUT == 2) {
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(&S.W3f[2 * k]);
        o[0] = fma2(w.x, pk2(t0, t0), o[0]);
        o[0] = fma2(w.y, pk2(t1, t1), o[0]);
    } else {
        const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(&S.W3f[4 * k]), w1 = *reinterpret_cast<const ulonglong2*>(&S.W3f[4 * k + 4]);
        o[0] = fma2(w0.x, pk2(t0, t0), o[0]);
        o[1] = fma2(w0.y, pk2(t0, t0), o[1]);
        o[0] = fma2(w1.x, pk2(t1, t1), o[0]);
        o[1] = fma2(w1.y, pk2(t1, t1), o[1]);
    }
}
template <int NOUT> __device__ __forceinline__ void final_chunk(const TcShared& S, const float* v, int cc, f32x2* o) {
#if RB_TANH_VARIANT == 5
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f32x2 t10, t32;
        float t0, t1, t2, t3;
        tanh_quad_packed(v + 4 * q, t10, t32);
        upk2(t10, t1, t0); upk2(t32, t3, t2);
        final_pair<NOUT>(S, 16 * cc + 4 * q, t0, t1, o);
        final_pair<NOUT>(S, 16 * cc + 4 * q + 2, t2, t3, o);
    }
#else
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float t0, t1;
#if RB_TANH_VARIANT == 4
        upk2(tanh_pair_packed(v[2 * q], v[2 * q + 1]), t1, t0);
#elif RB_TANH_VARIANT == 1
        tanh_pair_from_scaled(v[2 * q], v[2 * q + 1], t0, t1);
#else
        t0 = tanh_from_scaled(v[2 * q]); t1 = tanh_from_scaled(v[2 * q + 1]);
#endif
        final_pair<NOUT>(S, 16 * cc + 2 * q, t0, t1, o);
    }
#endif
}
template <int NOUT> __device__ __forceinline__ void final_epilogue(const TcShared& S, uint32_t taddr, f32x2* o) {
    float va[16], vb[16];
    tmem_ld_x16(taddr, va);
#pragma unroll 1
    for (int cc = 0; cc < 4; cc += 2) {
        tmem_ld_wait();
        tmem_ld_x16(taddr + 16 * (cc + 1), vb);
        final_chunk<NOUT>(S, va, cc, o);
        tmem_ld_wait();
        if (cc + 2 < 4) tmem_ld_x16(taddr + 16 * (cc + 2), va);
        final_chunk<NOUT>(S, vb, cc + 1, o);
    }
}

// Full policy evaluation for one tile.  Every ACTIVE thread of the tile must call this (it contains the tile barrier,
// `bar_threads` = 32 * active warps of the tile).  ob: this thread's 11-d observation.  pd: (mean0, mean1, logstd0,
// logstd1) or the four raw outputs.  issuer: true for exactly one (whole, active) warp of the tile; one elected lane issues the MMAs.
template <int NOUT, int TCOLS>
__device__ __forceinline__ void policy_tc_eval(TcShared& S, TcTile& T, int tile, int row, int bar_threads, bool issuer, const float* ob,
                                               float* pd, uint32_t& phase) {
    const uint32_t tmem1 = S.tmem_base + (uint32_t)tile * (uint32_t)TCOLS;       // layer-1 accumulator
    const uint32_t tmem2 = TCOLS >= 128 ? tmem1 + 64u : tmem1;                   // layer-2 accumulator (its own columns when they exist)
    const uint32_t lane_base = (uint32_t)(row & ~31) << 16;     // this warp's 32-lane TMEM window
    uint64_t* mbar = &S.mbar[tile];
    const uint32_t ah = smem_u32(T.A_hi), al = smem_u32(T.A_lo);
    const uint32_t idesc = make_idesc_bf16(128, 64);
    // ---- A1 = [clip((ob - mu) * inv_sd), 1, 0...] --------------------------------------------------------------
    {
        float z[16];
#pragma unroll
        for (int k = 0; k < 11; ++k) z[k] = fminf(5.f, fmaxf(-5.f, (ob[k] - S.mu[k]) * S.inv_sd[k]));
        z[11] = 1.f; z[12] = z[13] = z[14] = z[15] = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_pair(z[8 * c + 2 * q], z[8 * c + 2 * q + 1], h[q], l[q]);
            *reinterpret_cast<uint4*>(T.A_hi + c * 2048 + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(T.A_lo + c * 2048 + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
    // ---- layer 1: one K = 16 step (b1 rides in the ones column) ---------------------------------------------------
    fence_async_smem();
    fence_before_sync();
    tile_sync(1 + tile, bar_threads);
    if (issuer && elect_one_sync()) {                               // `issuer` is warp-uniform: one whole warp of the tile
        fence_after_sync();
        issue_bf16x3(tmem1, ah, al, smem_u32(S.B1_hi), smem_u32(S.B1_lo), 1, 64, idesc, 0);
        mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1u;
    fence_after_sync();
    // ---- layer-1 epilogue (h1 -> A tile of layer 2) with the layer-2 MMAs issued as soon as their K-steps are written ---------------
    auto issue_l2 = [&](int ks0, int ksteps, bool first, bool last) {
        if (issuer && elect_one_sync()) {
            fence_after_sync();
            if (first) {
                const uint64_t ones = make_smem_desc(smem_u32(S.ONES), 128 * 16, 128);
                mma_bf16(tmem2, ones, make_smem_desc(smem_u32(S.B2b_hi), 64 * 16, 128), idesc, 0);     // D = 1 * b2 (hi + lo)
                mma_bf16(tmem2, ones, make_smem_desc(smem_u32(S.B2b_lo), 64 * 16, 128), idesc, 1);
            }
            issue_bf16x3(tmem2, ah, al, smem_u32(S.B2_hi), smem_u32(S.B2_lo), ksteps, 64, idesc, 1, ks0);
            if (last) mma_commit(mbar);
        }
    };
    {
        const uint32_t taddr = tmem1 + lane_base;
        float va[16], vb[16];
        tmem_ld_x16(taddr, va);
#pragma unroll 1
        for (int cc = 0; cc < 4; cc += 2) {
            tmem_ld_wait();
            tmem_ld_x16(taddr + 16 * (cc + 1), vb);
            hidden_chunk(T, va, cc, row);
            tmem_ld_wait();
            if (cc + 2 < 4) tmem_ld_x16(taddr + 16 * (cc + 2), va);
            hidden_chunk(T, vb, cc + 1, row);
            if (TCOLS >= 128 || cc == 2) {                          // K-steps cc, cc + 1 of the layer-2 A tile are complete for this thread
                fence_async_smem();
                fence_before_sync();
                tile_sync(1 + tile, bar_threads);
                if (TCOLS >= 128) issue_l2(cc, 2, cc == 0, cc == 2);
                else issue_l2(0, 4, true, true);                     // shared accumulator: only after the whole layer-1 row has been read
            }
        }
    }
    mbar_wait(mbar, phase); phase ^= 1u;
    fence_after_sync();
    // ---- layer-2 epilogue + output layer, registers only -------------------------------------------------------------------------
    f32x2 o[2] = {pk2(S.b3[0], S.b3[1]), pk2(S.b3[2], S.b3[3])};
    final_epilogue<NOUT>(S, tmem2 + lane_base, o);
    upk2(o[0], pd[0], pd[1]);
    if (NOUT == 4) upk2(o[1], pd[2], pd[3]); else { pd[2] = S.logstd[0]; pd[3] = S.logstd[1]; }
    fence_before_sync();       // orders this TMEM read before the next evaluation's MMA (issued after the next tile barrier)
}

template <int NT> __device__ __forceinline__ void policy_tc_setup(TcShared& S, const float* params, int nout) {
    if (threadIdx.x < 32) tmem_alloc<tmem_cols<NT>()>(&S.tmem_base);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < MAX_TILES; ++t) mbar_init(&S.mbar[t], 1);
        fence_mbar_init();
    }
    policy_tc_load_weights(S, params, nout);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}
// same, with the split weights coming from a prebuilt global image (k_policy_tc_build_image) through one TMA bulk copy
template <int NT> __device__ __forceinline__ void policy_tc_setup_from_image(TcShared& S, const void* img) {
    if (threadIdx.x < 32) tmem_alloc<tmem_cols<NT>()>(&S.tmem_base);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < MAX_TILES; ++t) mbar_init(&S.mbar[t], 1);
        mbar_init(&S.mbar_load, 1);
        fence_mbar_init();
        mbar_expect_tx(&S.mbar_load, TC_IMAGE_BYTES);
        bulk_g2s(&S, img, TC_IMAGE_BYTES, &S.mbar_load);
    }
    __syncthreads();                      // barrier inits visible before anyone waits
    mbar_wait(&S.mbar_load, 0);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}
template <int NT> __device__ __forceinline__ void policy_tc_teardown(TcShared& S) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<tmem_cols<NT>()>(S.tmem_base);
}

// Standalone forward: every CTA strides over groups of NT tiles (all 4 warps of every tile take part; rows >= n are zeros)
template <int NOUT, int NT>
__global__ void __launch_bounds__(NT* TILE, 1) k_policy_fwd_tc(const float* __restrict__ params, const float* __restrict__ obs, int64_t n,
                                                                float4* __restrict__ pd_out) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcShared& S = *reinterpret_cast<TcShared*>(smem_raw);
    TcTile* tiles = reinterpret_cast<TcTile*>(smem_raw + sizeof(TcShared));
    policy_tc_setup<NT>(S, params, NOUT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tile = warp >> 2, row = threadIdx.x & (TILE - 1);
    TcTile& T = tiles[tile];
    float* strip = reinterpret_cast<float*>(T.A_hi + STRIP_OFF) + (warp & 3) * 32 * OBS;
    uint32_t phase = 0;
    const int64_t ngroups = (n + (int64_t)NT * TILE - 1) / ((int64_t)NT * TILE);
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int64_t i = g * NT * TILE + threadIdx.x;
        const int64_t row0 = i - lane;
        const int nvalid = (int)max((int64_t)0, min((int64_t)32, n - row0));
        float ob[OBS], pd[4];
        warp_load_rows<OBS>(obs, row0, nvalid, ob, strip, lane);
        __syncwarp();
        policy_tc_eval<NOUT, tile_cols<NT>()>(S, T, tile, row, TILE, row < 32, ob, pd, phase);
        if (i < n) pd_out[i] = make_float4(pd[0], pd[1], pd[2], pd[3]);
    }
    policy_tc_teardown<NT>(S);
}

__global__ void k_policy_tc_build_image(const float* __restrict__ params, int nout, TcShared* img) { policy_tc_load_weights(*img, params, nout); }

// DAgger observe, fused: for every env the observation, the teacher label t = teacher(ob) (mlp_train.py:165-167) and the student
// input row (MLP student: [dropout(ob), prev teacher pdflat, prev recorded reward], zeros at the first step of an episode --
// mlp_train.py:50-52, dataset.py:118-143; 2x64 student: x = ob).  One pass over the env state, teacher on tcgen05.
template <int KIND, int NT>
__global__ void __launch_bounds__(NT* TILE, 1) k_dagger_observe_tc(int64_t n, const float4* __restrict__ qv, const float4* __restrict__ tp,
                                                                    const uint4* __restrict__ ctr, const void* __restrict__ teacher_img,
                                                                    const float4* __restrict__ prev_t, const float* __restrict__ prev_rec_rew,
                                                                    float keep_prob, uint32_t k0, uint32_t k1, uint32_t offset, uint32_t iteration,
                                                                    const uint32_t* __restrict__ clock, float* __restrict__ obs_out,
                                                                    float4* __restrict__ t_out, float* __restrict__ x_out, float* __restrict__ x_act_out) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    if (clock) iteration = clock[0];          // device-side step clock (CUDA-graph replay)
    TcShared& S = *reinterpret_cast<TcShared*>(smem_raw);
    TcTile* tiles = reinterpret_cast<TcTile*>(smem_raw + sizeof(TcShared));
    policy_tc_setup_from_image<NT>(S, teacher_img);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tile = warp >> 2, row = threadIdx.x & (TILE - 1);
    TcTile& T = tiles[tile];
    float* strip = reinterpret_cast<float*>(T.A_hi + STRIP_OFF) + (warp & 3) * 32 * OBS;
    uint32_t phase = 0;
    const int64_t ngroups = (n + (int64_t)NT * TILE - 1) / ((int64_t)NT * TILE);
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int64_t i = g * NT * TILE + threadIdx.x;
        const int64_t row0 = i - lane;
        const int nvalid = (int)max((int64_t)0, min((int64_t)32, n - row0));
        const bool valid = i < n;
        float ob[OBS], pd[4];
        uint32_t step = 1u;
        if (valid) {
            const EnvState e = load_state(qv, tp, ctr, i);
            observe(e, ob);
            step = (uint32_t)e.step;
        } else {
#pragma unroll
            for (int k = 0; k < OBS; ++k) ob[k] = 0.f;
        }
        if (nvalid > 0) {
            warp_store_rows<OBS>(obs_out, row0, nvalid, ob, strip, lane);
            if (KIND == RB_STUDENT_POLICY64 && x_out != obs_out) warp_store_rows<OBS>(x_out, row0, nvalid, ob, strip, lane);
        }
        __syncwarp();
        policy_tc_eval<2, tile_cols<NT>()>(S, T, tile, row, TILE, row < 32, ob, pd, phase);
        if (valid) {
            t_out[i] = make_float4(pd[0], pd[1], pd[2], pd[3]);
            if (KIND == RB_STUDENT_MLP) {
                const bool first = step == 0u;   // first record of an episode: prev / prew are zeros (dataset.py:151-164)
                const float4 pp = first ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(prev_t + i);
                const float pr = first ? 0.f : __ldg(prev_rec_rew + i);
                float4 o[4];
                mlp_input_row(ob, keep_prob, k0, k1, offset + (uint32_t)i, iteration, pp, pr, o);
                float4* xr = reinterpret_cast<float4*>(x_out) + i * 4;
                xr[0] = o[0]; xr[1] = o[1]; xr[2] = o[2]; xr[3] = o[3];
                if (x_act_out) {         // the same row without observation dropout: what the student ACTS on (mlp_train.py:171-186, keep_prob 1)
                    float4* xa = reinterpret_cast<float4*>(x_act_out) + i * 4;
                    xa[0] = make_float4(ob[0], ob[1], ob[2], ob[3]); xa[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
                    xa[2] = make_float4(ob[8], ob[9], ob[10], pp.x); xa[3] = make_float4(pp.y, pp.z, pp.w, pr);
                }
            }
        }
    }
    policy_tc_teardown<NT>(S);
}

// Fused policy-in-the-loop rollout (teacher warm-up loop, mlp_train.py:120-139).  Envs are dealt to CTAs in warp units.
template <int NOUT, int NT>
__global__ void __launch_bounds__(NT* TILE, 1) k_rollout_policy_tc(int64_t n, float4* qv, float4* tp, uint4* ctr, const float* __restrict__ params,
                                                                    int T, float* __restrict__ obs_buf, float4* __restrict__ pd_buf,
                                                                    float* __restrict__ rew_buf, uint8_t* __restrict__ done_buf, uint32_t k0,
                                                                    uint32_t k1, uint32_t offset, uint32_t stagger_ns, uint32_t* prog_counters,
                                                                    uint32_t* prog_flags, int prog_slab_len, uint32_t prog_epoch,
                                                                    uint64_t* __restrict__ done_mask, float* __restrict__ return_sum) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcShared& S = *reinterpret_cast<TcShared*>(smem_raw);
    TcTile* tiles = reinterpret_cast<TcTile*>(smem_raw + sizeof(TcShared));
    // balanced contiguous split of the ceil(n/32) warp units over the grid
    const int64_t units = (n + 31) >> 5, per = units / gridDim.x, rem = units % gridDim.x;
    const int64_t unit0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
    const int nunits = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));       // <= 4 * NT (host guarantees)
    if (nunits == 0) return;                                                  // CTA-uniform
    policy_tc_setup<NT>(S, params, NOUT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tile = warp >> 2, row = threadIdx.x & (TILE - 1);
    if (warp < nunits) {
        TcTile& Tl = tiles[tile];
        const int bar_threads = 32 * min(4, nunits - 4 * tile);
        float* strip = reinterpret_cast<float*>(Tl.A_hi + STRIP_OFF) + (warp & 3) * 32 * OBS;
        const int64_t row0 = (unit0 + warp) << 5, i = row0 + lane;
        const int nvalid = (int)min((int64_t)32, n - row0);
        const bool valid = i < n;
        const uint32_t gid = offset + (uint32_t)i;
        EnvState e = valid ? load_state(qv, tp, ctr, i) : zero_state();
        uint32_t phase = 0;
        // The tiles of a CTA alternate between an XU-bound phase (tanh epilogues: 2 MUFU per element) and an issue-bound phase
        // (physics).  Started together they stay in lock step and fight for the same pipe; a one-time offset of a fraction of the
        // step period lets one tile's epilogue overlap another tile's physics.
        for (uint32_t w = 0; w < (uint32_t)tile * stagger_ns; w += 1000u) __nanosleep(1000u);
        int next_mark = prog_slab_len > 0 ? min(prog_slab_len, T) : -1;      // step count at which the next time slab is complete (-1: no reporting)
        uint64_t dmask = 0ull;                                               // bit t = this env finished an episode at step t of the chunk
        float rsum = 0.f;                                                    // sum of this env's rewards over the chunk, in step order
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            float ob[OBS], pd[4];
            observe(e, ob);
            if (obs_buf) warp_store_rows<OBS>(obs_buf, (int64_t)t * n + row0, nvalid, ob, strip, lane);
            __syncwarp();
            policy_tc_eval<NOUT, tile_cols<NT>()>(S, Tl, tile, row, bar_threads, row < 32, ob, pd, phase);
            bool d;
            const float rw = step_env(e, pd[0], pd[1], k0, k1, gid, d);
            dmask |= (uint64_t)(d ? 1u : 0u) << (t & 63);
            rsum = __fadd_rn(rsum, rw);
            const int64_t r = (int64_t)t * n + i;
            if (valid) {
                if (pd_buf) pd_buf[r] = make_float4(pd[0], pd[1], pd[2], pd[3]);
                if (rew_buf) rew_buf[r] = rw;
                if (done_buf) done_buf[r] = d ? 1 : 0;
            }
            if (t + 1 == next_mark) {                                                     // warp-uniform: a time slab of the buffer is complete
                next_mark = min(next_mark + prog_slab_len, T);
                // release at GPU scope by every lane, then the count; the ONE warp that completes the count issues the system-scope fence below, and
                // fence cumulativity (PTX memory model) carries every counted warp's rows with it.  A system-scope fence here, per warp, has to
                // drain that warp's posted PCIe reward stores and cost 0.07 ms per 0.30 ms call (e2e 8.9e9 -> 7.4e9 env-steps/s).
                __threadfence();
                __syncwarp();                                                             // ... for every lane of the warp, before the count
                if (lane == 0) {
                    const int slab = t / prog_slab_len;
                    if (atomicAdd(prog_counters + slab, 1u) + 1u == (uint32_t)units) {      // counters are zeroed by the host call
                        __threadfence_system();                                           // kernel -> host -> copy engine: system-scope ordering
                        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(prog_flags + slab), "r"(prog_epoch) : "memory");
                    }
                }
            }
        }
        if (valid) {
            store_state(qv, tp, ctr, i, e);
            if (done_mask) done_mask[i] = dmask;                              // 8 B per env per chunk instead of T bytes (256 contiguous bytes per warp)
            if (return_sum) return_sum[i] = rsum;
        }
    }
    policy_tc_teardown<NT>(S);
}

template <typename K> static int set_smem_attr(K kern, size_t bytes) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RB_OK;
}

static int sm_count_current(int* sms) {
    int device = 0;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
    return RB_OK;
}

constexpr int FWD_NT = 2;       // standalone forward: 2 tiles per CTA (97 KB) -> 2 CTAs per SM
constexpr int ROLLOUT_NT = 4;   // fused rollout: 4 tiles = 16 warps per CTA, one CTA per SM
constexpr int ROLLOUT_STAGGER_NS = 1000;   // start offset between consecutive tiles of a CTA (step period 5.9 us; swept 0 .. 4000 ns in round 2: 1.115 / 1.120 / 1.112 / 1.093e10 env-steps/s at 0 / 1000 / 2000 / 4000)

int policy_fwd_tc(const float* params, int nout, const float* obs, int64_t n, float* pd, cudaStream_t s) {
    int sms = 148;
    int rc = sm_count_current(&sms);
    if (rc) return rc;
    const int64_t ngroups = (n + FWD_NT * TILE - 1) / (FWD_NT * TILE);
    const unsigned grid = (unsigned)min((int64_t)sms * 2, ngroups);
    const size_t smem = tc_smem_bytes<FWD_NT>();
    if (nout == 2) {
        rc = set_smem_attr(k_policy_fwd_tc<2, FWD_NT>, smem);
        if (rc) return rc;
        k_policy_fwd_tc<2, FWD_NT><<<grid, FWD_NT * TILE, smem, s>>>(params, obs, n, (float4*)pd);
    } else {
        rc = set_smem_attr(k_policy_fwd_tc<4, FWD_NT>, smem);
        if (rc) return rc;
        k_policy_fwd_tc<4, FWD_NT><<<grid, FWD_NT * TILE, smem, s>>>(params, obs, n, (float4*)pd);
    }
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

size_t policy_tc_image_bytes() { return TC_IMAGE_BYTES; }

int policy_tc_build_image(const float* params, int nout, void* img, cudaStream_t s) {
    k_policy_tc_build_image<<<1, 256, 0, s>>>(params, nout, (TcShared*)img);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int dagger_observe_tc(rb_env* e, const void* teacher_img, int student_kind, float keep_prob, const float4* prev_t, const float* prev_rec_rew,
                      uint32_t iteration, const uint32_t* clock, float* obs, float* t_pd, float* x, float* x_act, cudaStream_t s) {
    constexpr int NT = FWD_NT;
    const int64_t ngroups = (e->n + NT * TILE - 1) / (NT * TILE);
    const unsigned grid = (unsigned)min((int64_t)e->sm_count * 2, ngroups);
    const size_t smem = tc_smem_bytes<NT>();
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    if (student_kind == RB_STUDENT_MLP) {
        int rc = set_smem_attr(k_dagger_observe_tc<RB_STUDENT_MLP, NT>, smem);
        if (rc) return rc;
        k_dagger_observe_tc<RB_STUDENT_MLP, NT><<<grid, NT * TILE, smem, s>>>(e->n, e->qv, e->tp, e->ctr, teacher_img, prev_t, prev_rec_rew, keep_prob, k0,
                                                                            k1, e->offset, iteration, clock, obs, (float4*)t_pd, x, x_act);
    } else {
        int rc = set_smem_attr(k_dagger_observe_tc<RB_STUDENT_POLICY64, NT>, smem);
        if (rc) return rc;
        k_dagger_observe_tc<RB_STUDENT_POLICY64, NT><<<grid, NT * TILE, smem, s>>>(e->n, e->qv, e->tp, e->ctr, teacher_img, prev_t, prev_rec_rew,
                                                                                 keep_prob, k0, k1, e->offset, iteration, clock, obs, (float4*)t_pd, x, x_act);
    }
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

template <int NOUT, int NT>
static int launch_rollout_tc(rb_env* e, const float* params, int T, float* obs_buf, float* pd_buf, float* rew_buf, uint8_t* done_buf, cudaStream_t s) {
    const int64_t units = (e->n + 31) / 32;
    const int64_t per_round = (int64_t)4 * NT * e->sm_count;
    const int64_t rounds = (units + per_round - 1) / per_round;
    const unsigned grid = (unsigned)min(units, rounds * e->sm_count);        // units / grid <= 4 * NT
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    const size_t smem = tc_smem_bytes<NT>();
    int rc = set_smem_attr(k_rollout_policy_tc<NOUT, NT>, smem);
    if (rc) return rc;
    const int stagger = ROLLOUT_STAGGER_NS;
    k_rollout_policy_tc<NOUT, NT><<<grid, NT * TILE, smem, s>>>(e->n, e->qv, e->tp, e->ctr, params, T, obs_buf, (float4*)pd_buf, rew_buf, done_buf, k0, k1,
                                                                e->offset, (uint32_t)stagger, e->prog_counters, e->prog_flags_dev,
                                                                e->prog_counters ? e->prog_slab_len : 0, e->prog_epoch, e->done_mask_out, e->return_sum_out);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rollout_policy_tc(rb_env* e, const float* params, int nout, int T, float* obs_buf, float* pd_buf, float* rew_buf, uint8_t* done_buf,
                      cudaStream_t s) {
    if (nout == 2) return launch_rollout_tc<2, ROLLOUT_NT>(e, params, T, obs_buf, pd_buf, rew_buf, done_buf, s);
    return launch_rollout_tc<4, ROLLOUT_NT>(e, params, T, obs_buf, pd_buf, rew_buf, done_buf, s);
}

}  // namespace rb
