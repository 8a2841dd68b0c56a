// RB_MODE_TC: the baselines MlpPolicy (11 -> 64 tanh -> 64 tanh -> nout) on the 5th-gen tensor cores.
// Replaces sess.run((pi.pd.mean, pi.pd.flat)) (/root/reference src/distilation/mlp_train.py:123-125,165-167; network
// teacher.py:12-16) and, fused with the env step, the teacher warm-up loop mlp_train.py:120-139.
//
// One CTA = 128 threads = 128 envs/samples = the 128 TMEM lanes of one M=128 accumulator tile; thread r owns row r
// end to end (its env state lives in its registers, its accumulator row comes back through tcgen05.ld 32x32b).
// Per policy evaluation the CTA runs three dependent GEMMs  [128 x K] * [K x N]:
//     L1: K = 16 (11 obs + a ones column that carries b1 + zero pad), N = 64  -> TMEM cols   0.. 63
//     L2: K = 64, N = 64                                                      -> TMEM cols  64..127
//     L3: K = 64, N = 16 (nout padded)                                        -> TMEM cols   0.. 15
// Operands are bf16 hi/lo splits of the fp32 values ("bf16x3": A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, fp32 accumulate in
// TMEM), which keeps the result within ~1e-5 of the fp32 network while running on tcgen05.  Weights are split once per
// CTA into shared memory; activations are split in the epilogue (bias + MUFU tanh) and written straight into the next
// layer's A tile (no-swizzle K-major layout, see tc_common.cuh).  One elected thread issues the MMAs; completion is
// tracked with tcgen05.commit -> mbarrier.
#include "common.cuh"
#include "physics.cuh"
#include "tc_common.cuh"

namespace rb {

using namespace tc;

constexpr int TCB = 128;          // threads per CTA == rows per tile
constexpr int N_TERMS = 3;        // bf16x3

struct __align__(128) PolicyTcSmem {
    uint8_t A_hi[128 * 64 * 2];   // [8 chunks][128 rows][16 B]; the L1 tile (2 chunks) aliases its head; tail doubles as obs strips
    uint8_t A_lo[128 * 64 * 2];
    uint8_t B2_hi[64 * 64 * 2];
    uint8_t B2_lo[64 * 64 * 2];
    uint8_t B1_hi[64 * 16 * 2];
    uint8_t B1_lo[64 * 16 * 2];
    uint8_t B3_hi[16 * 64 * 2];
    uint8_t B3_lo[16 * 64 * 2];
    float b2[64];
    float b3[4];
    float mu[12];
    float inv_sd[12];
    float logstd[4];
    uint64_t mbar;
    uint32_t tmem_base;
};
constexpr uint32_t STRIP_OFF = 8192;   // obs staging strips live in A_hi[8192 .. 8192 + 4*1408)

// split W (fp32, row-major [K][N] in global) into the K-major B tiles (element (n,k) = W[k][n])
__device__ inline void policy_tc_load_weights(PolicyTcSmem& S, const float* __restrict__ p, int nout) {
    const PolicyOffsets o = policy_offsets(nout);
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 16; i += TCB) {          // B1: n = i % 64, k = i / 64 ; k == 11 carries the bias
        const int n = i & 63, k = i >> 6;
        const float w = k < 11 ? __ldg(p + o.W1 + k * 64 + n) : (k == 11 ? __ldg(p + o.b1 + n) : 0.f);
        uint16_t h, l;
        split_scalar(w, h, l);
        *reinterpret_cast<uint16_t*>(S.B1_hi + tile_off(n, k, 64)) = h;
        *reinterpret_cast<uint16_t*>(S.B1_lo + tile_off(n, k, 64)) = l;
    }
    for (int i = tid; i < 64 * 64; i += TCB) {
        const int n = i & 63, k = i >> 6;
        uint16_t h, l;
        split_scalar(__ldg(p + o.W2 + k * 64 + n), h, l);
        *reinterpret_cast<uint16_t*>(S.B2_hi + tile_off(n, k, 64)) = h;
        *reinterpret_cast<uint16_t*>(S.B2_lo + tile_off(n, k, 64)) = l;
    }
    for (int i = tid; i < 16 * 64; i += TCB) {
        const int n = i & 15, k = i >> 4;
        uint16_t h, l;
        split_scalar(n < nout ? __ldg(p + o.W3 + k * nout + n) : 0.f, h, l);
        *reinterpret_cast<uint16_t*>(S.B3_hi + tile_off(n, k, 16)) = h;
        *reinterpret_cast<uint16_t*>(S.B3_lo + tile_off(n, k, 16)) = l;
    }
    if (tid < 64) S.b2[tid] = __ldg(p + o.b2 + tid);
    if (tid < 4) {
        S.b3[tid] = tid < nout ? __ldg(p + o.b3 + tid) : 0.f;
        S.logstd[tid] = tid < 2 ? __ldg(p + o.logstd + tid) : 0.f;
    }
    if (tid < 12) {
        S.mu[tid] = tid < 11 ? __ldg(p + o.mu + tid) : 0.f;
        S.inv_sd[tid] = tid < 11 ? 1.0f / __ldg(p + o.sd + tid) : 0.f;
    }
}

// issue one layer: D[tmem_d] = sum over k-steps and hi/lo terms of A * B^T.  Called by ONE thread.
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* b_hi,
                                            const uint8_t* b_lo, int ksteps, int n_rows_b, uint32_t idesc) {
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)n_rows_b * 16;
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t dah = make_smem_desc(ah + ks * 2 * a_lbo, a_lbo, 128), dal = make_smem_desc(al + ks * 2 * a_lbo, a_lbo, 128);
        const uint64_t dbh = make_smem_desc(bh + ks * 2 * b_lbo, b_lbo, 128), dbl = make_smem_desc(bl + ks * 2 * b_lbo, b_lbo, 128);
        mma_bf16(tmem_d, dah, dbh, idesc, acc);
        acc = 1;
        if (N_TERMS >= 2) mma_bf16(tmem_d, dal, dbh, idesc, 1);
        if (N_TERMS >= 3) mma_bf16(tmem_d, dah, dbl, idesc, 1);
        if (N_TERMS >= 4) mma_bf16(tmem_d, dal, dbl, idesc, 1);
    }
}

// epilogue of a 64-wide hidden layer: TMEM row -> (+bias) -> tanh -> bf16 hi/lo -> this thread's row of the next A tile
__device__ __forceinline__ void hidden_epilogue(PolicyTcSmem& S, uint32_t taddr, const float* bias, int row) {
    float v[64];
    tmem_ld_x16(taddr, v);
    tmem_ld_x16(taddr + 16, v + 16);
    tmem_ld_x16(taddr + 32, v + 32);
    tmem_ld_x16(taddr + 48, v + 48);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float a = v[8 * c + 2 * q], b = v[8 * c + 2 * q + 1];
            if (bias) { a += bias[8 * c + 2 * q]; b += bias[8 * c + 2 * q + 1]; }
            split_pair(tanh_mufu(a), tanh_mufu(b), h[q], l[q]);
        }
        *reinterpret_cast<uint4*>(S.A_hi + c * 2048 + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(S.A_lo + c * 2048 + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// Full policy evaluation for the CTA's 128 rows.  Every thread of the CTA must call this (it contains CTA barriers).
// ob: this thread's 11-d observation.  pd: (mean0, mean1, logstd0, logstd1) or the four raw outputs.
template <int NOUT>
__device__ __forceinline__ void policy_tc_eval(PolicyTcSmem& S, const float* ob, float* pd, uint32_t& phase) {
    const int row = threadIdx.x;
    const uint32_t tmem = S.tmem_base;
    const uint32_t lane_base = (uint32_t)(row & ~31) << 16;     // this warp's 32-lane TMEM window
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
            *reinterpret_cast<uint4*>(S.A_hi + c * 2048 + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(S.A_lo + c * 2048 + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        issue_layer(tmem, S.A_hi, S.A_lo, S.B1_hi, S.B1_lo, 1, 64, make_idesc_bf16(128, 64));
        mma_commit(&S.mbar);
    }
    mbar_wait(&S.mbar, phase); phase ^= 1u;
    fence_after_sync();
    hidden_epilogue(S, tmem + lane_base, nullptr, row);           // b1 rides in the ones column
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        issue_layer(tmem + 64, S.A_hi, S.A_lo, S.B2_hi, S.B2_lo, 4, 64, make_idesc_bf16(128, 64));
        mma_commit(&S.mbar);
    }
    mbar_wait(&S.mbar, phase); phase ^= 1u;
    fence_after_sync();
    hidden_epilogue(S, tmem + lane_base + 64, S.b2, row);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        issue_layer(tmem, S.A_hi, S.A_lo, S.B3_hi, S.B3_lo, 4, 16, make_idesc_bf16(128, 16));
        mma_commit(&S.mbar);
    }
    mbar_wait(&S.mbar, phase); phase ^= 1u;
    fence_after_sync();
    float o[4];
    tmem_ld_x4(tmem + lane_base, o);
    tmem_ld_wait();
    pd[0] = o[0] + S.b3[0];
    pd[1] = o[1] + S.b3[1];
    if (NOUT == 4) { pd[2] = o[2] + S.b3[2]; pd[3] = o[3] + S.b3[3]; } else { pd[2] = S.logstd[0]; pd[3] = S.logstd[1]; }
    fence_before_sync();       // orders this TMEM read before the next evaluation's MMA (issued after the next barrier)
}

__device__ __forceinline__ void policy_tc_setup(PolicyTcSmem& S, const float* params, int nout) {
    if (threadIdx.x < 32) tmem_alloc<128>(&S.tmem_base);
    if (threadIdx.x == 0) { mbar_init(&S.mbar, 1); fence_mbar_init(); }
    policy_tc_load_weights(S, params, nout);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}
__device__ __forceinline__ void policy_tc_teardown(PolicyTcSmem& S) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<128>(S.tmem_base);
}

template <int NOUT>
__global__ void __launch_bounds__(TCB) k_policy_fwd_tc(const float* __restrict__ params, const float* __restrict__ obs, int64_t n,
                                                       float4* __restrict__ pd_out) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    PolicyTcSmem& S = *reinterpret_cast<PolicyTcSmem*>(smem_raw);
    policy_tc_setup(S, params, NOUT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* strip = reinterpret_cast<float*>(S.A_hi + STRIP_OFF) + warp * 32 * OBS;
    uint32_t phase = 0;
    const int64_t ntiles = (n + TCB - 1) / TCB;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i = tile * TCB + threadIdx.x;
        const int64_t row0 = i - lane;
        const int nvalid = (int)max((int64_t)0, min((int64_t)32, n - row0));
        float ob[OBS], pd[4];
        warp_load_rows<OBS>(obs, row0, nvalid, ob, strip, lane);
        __syncwarp();
        policy_tc_eval<NOUT>(S, ob, pd, phase);
        if (i < n) pd_out[i] = make_float4(pd[0], pd[1], pd[2], pd[3]);
    }
    policy_tc_teardown(S);
}

template <int NOUT>
__global__ void __launch_bounds__(TCB) k_rollout_policy_tc(int64_t n, float4* qv, float4* tp, uint2* ctr, const float* __restrict__ params,
                                                           int T, float* __restrict__ obs_buf, float4* __restrict__ pd_buf,
                                                           float* __restrict__ rew_buf, uint8_t* __restrict__ done_buf, uint32_t k0,
                                                           uint32_t k1, uint32_t offset) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    PolicyTcSmem& S = *reinterpret_cast<PolicyTcSmem*>(smem_raw);
    policy_tc_setup(S, params, NOUT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* strip = reinterpret_cast<float*>(S.A_hi + STRIP_OFF) + warp * 32 * OBS;
    const int64_t i = (int64_t)blockIdx.x * TCB + threadIdx.x;
    const int64_t row0 = i - lane;
    const int nvalid = (int)max((int64_t)0, min((int64_t)32, n - row0));
    const bool valid = i < n;
    const uint32_t gid = offset + (uint32_t)i;
    EnvState<float> e;
    if (valid) {
        const float4 a = qv[i], b = tp[i];
        const uint2 c = ctr[i];
        e.q0 = a.x; e.q1 = a.y; e.v0 = a.z; e.v1 = a.w; e.tx = b.x; e.ty = b.y; e.px = b.z; e.py = b.w; e.step = (int)c.x; e.episode = c.y;
    } else { e.q0 = e.q1 = e.v0 = e.v1 = e.tx = e.ty = e.px = e.py = 0.f; e.step = 0; e.episode = 0; }
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
        float ob[OBS], pd[4];
        observe(e, ob);
        if (obs_buf && nvalid > 0) warp_store_rows<OBS>(obs_buf, (int64_t)t * n + row0, nvalid, ob, strip, lane);
        __syncwarp();
        policy_tc_eval<NOUT>(S, ob, pd, phase);
        bool d;
        const float rw = step_env(e, pd[0], pd[1], k0, k1, gid, d);
        const int64_t row = (int64_t)t * n + i;
        if (valid) {
            if (pd_buf) pd_buf[row] = make_float4(pd[0], pd[1], pd[2], pd[3]);
            if (rew_buf) rew_buf[row] = rw;
            if (done_buf) done_buf[row] = d ? 1 : 0;
        }
    }
    if (valid) {
        qv[i] = make_float4(e.q0, e.q1, e.v0, e.v1);
        tp[i] = make_float4(e.tx, e.ty, e.px, e.py);
        ctr[i] = make_uint2((uint32_t)e.step, e.episode);
    }
    policy_tc_teardown(S);
}

template <typename K> static int set_smem_attr(K kern) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PolicyTcSmem)));
    return RB_OK;
}

int policy_fwd_tc(const float* params, int nout, const float* obs, int64_t n, float* pd, cudaStream_t s) {
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t ntiles = (n + TCB - 1) / TCB;
    const unsigned grid = (unsigned)min((int64_t)sms * 3, ntiles);
    if (nout == 2) {
        int rc = set_smem_attr(k_policy_fwd_tc<2>);
        if (rc) return rc;
        k_policy_fwd_tc<2><<<grid, TCB, sizeof(PolicyTcSmem), s>>>(params, obs, n, (float4*)pd);
    } else {
        int rc = set_smem_attr(k_policy_fwd_tc<4>);
        if (rc) return rc;
        k_policy_fwd_tc<4><<<grid, TCB, sizeof(PolicyTcSmem), s>>>(params, obs, n, (float4*)pd);
    }
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rollout_policy_tc(rb_env* e, const float* params, int nout, int T, float* obs_buf, float* pd_buf, float* rew_buf, uint8_t* done_buf,
                      cudaStream_t s) {
    const unsigned grid = (unsigned)((e->n + TCB - 1) / TCB);
    const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
    if (nout == 2) {
        int rc = set_smem_attr(k_rollout_policy_tc<2>);
        if (rc) return rc;
        k_rollout_policy_tc<2><<<grid, TCB, sizeof(PolicyTcSmem), s>>>(e->n, e->qv, e->tp, e->ctr, params, T, obs_buf, (float4*)pd_buf, rew_buf,
                                                                        done_buf, k0, k1, e->offset);
    } else {
        int rc = set_smem_attr(k_rollout_policy_tc<4>);
        if (rc) return rc;
        k_rollout_policy_tc<4><<<grid, TCB, sizeof(PolicyTcSmem), s>>>(e->n, e->qv, e->tp, e->ctr, params, T, obs_buf, (float4*)pd_buf, rew_buf,
                                                                        done_buf, k0, k1, e->offset);
    }
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // namespace rb
