// LSTM student recurrence (/root/reference src/distilation/student_nn.py:23,41: tf.contrib.rnn.LSTMCell(200), forget_bias 1, unrolled
// over the T = 10 steps of a window; trained by back-propagation through time, lstm_train.py:74-79) as TWO persistent launches: one for
// the forward recurrence, one for BPTT.  They replace 20 dependent GEMM launches + 20 element-wise cell launches + split-K reductions.
//
// Decomposition.  Window rows are independent, time steps are not, so a *row block* of 128 windows (one M = 128 tcgen05 accumulator)
// runs all T steps inside one thread-block CLUSTER of 8 CTAs without any grid-wide synchronisation.  The 200 LSTM units are dealt to
// the 8 CTAs (25 each); CTA j keeps, resident in shared memory for the whole launch, the bf16 hi/lo image of ITS slice of W_l:
//   forward : z[128 x 100] = [x_t | m_{t-1}] [128 x 256] * W_l[:, cols of its 25 units x 4 gates]          (N = 112, K = 256)
//             -> gates, c (registers, carried over the steps), m -> global (hh_t for the heads, xh_{t+1} for the next step)
//   backward: dz of its 25 units from dm, dc (registers) and the saved gates -> A tile [128 x 112]; the recurrent / input gradient
//             d[x | m_prev] = dz W_l^T is a sum over ALL units, so every CTA multiplies its K-slice, dz_j [128 x 112] * W_l^T[112 x 256],
//             into a PARTIAL [128 x 256] tile (N = 256), written column-major to global; at the next step the owner of a unit adds
//             the 8 partials in CTA order (deterministic).
// The only exchange between the CTAs of a cluster is m_t (forward) / the partial tiles (backward), through global memory (L2), ordered
// by one barrier.cluster (release / acquire) per step.  Operands follow the bf16x3 scheme of the other tensor-core kernels
// (A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM).
#include "common.cuh"
#include "lstm_recur.cuh"
#include "tc_common.cuh"

namespace rb {

using namespace tc;

namespace {

constexpr int RT = 10, RU = 200, RG = 800, RLD = 256, RX = 43, RKX = 243;
constexpr int RC_THREADS = 512, RC_SPLIT = 8, RC_UN = 25, RC_NL = 100, RC_NP = 112, RC_UPT = 7;
constexpr uint32_t IMG_HALF = 57344, IMG_SLICE = 2 * IMG_HALF;          // one bf16 half / hi + lo of one CTA's weight slice
constexpr int DZ_LD = 101;                                              // padded row of the dz staging tile (conflict-free)

struct __align__(128) FwdSmem {
    uint8_t W[IMG_SLICE];            // hi | lo : [32 k-chunks][112 rows (n = 4 u + gate)][16 B]
    uint8_t A[2][32768];             // two k-quarter buffers, each hi (16 KB) | lo (16 KB): [8 k-chunks][128 rows][16 B]
    float mst[128 * RC_UN];          // m (last step also c) staging for coalesced row stores
    float bias[RC_NP];
    uint64_t bar_w, bar_buf[2], bar_done;
    uint32_t tmem_base;
};
struct __align__(128) BwdSmem {
    uint8_t Bm[IMG_SLICE];           // hi | lo : [14 k-chunks][256 rows (xh column)][16 B]
    uint8_t A[2 * 28672];            // hi | lo : [14 k-chunks][128 rows][16 B]  (k = 4 u + gate, 100 used)
    float S[128 * DZ_LD];            // dL/dm tile of the step (first 3200 floats), then the dz staging tile
    uint64_t bar_w, bar_done;
    uint32_t tmem_base;
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float sigmoid_mufu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
}
// column of W_l / z that local gate column n = 4 u + g of CTA j stands for
__device__ __forceinline__ int gate_col(int j, int n) { return (n & 3) * RU + RC_UN * j + (n >> 2); }
// private (coalesced) index of the value a thread keeps for (step slot, row block, CTA, unit slot)
__device__ __forceinline__ size_t pidx(int slot, int nrb, int rb, int j, int i, int tid) {
    return ((((size_t)slot * nrb + rb) * RC_SPLIT + j) * RC_UPT + i) * RC_THREADS + tid;
}

// bf16 hi/lo images of W_l for both kernels: slice j, forward tile (rows n, K = xh column k) and backward tile (rows = xh column, K = n)
__global__ void k_lstm_recur_images(const float* __restrict__ W_l, uint8_t* __restrict__ img_f, uint8_t* __restrict__ img_b) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= RC_SPLIT * RLD * RC_NP) return;
    const int j = idx / (RLD * RC_NP), rem = idx - j * (RLD * RC_NP), k = rem / RC_NP, n = rem - k * RC_NP;
    const float w = (n < RC_NL && k < RKX) ? __ldg(W_l + (size_t)k * RG + gate_col(j, n)) : 0.f;
    uint16_t h, l;
    split_scalar(w, h, l);
    uint8_t* f = img_f + (size_t)j * IMG_SLICE + tile_off(n, k, RC_NP);
    *reinterpret_cast<uint16_t*>(f) = h;
    *reinterpret_cast<uint16_t*>(f + IMG_HALF) = l;
    uint8_t* b = img_b + (size_t)j * IMG_SLICE + tile_off(k, n, RLD);
    *reinterpret_cast<uint16_t*>(b) = h;
    *reinterpret_cast<uint16_t*>(b + IMG_HALF) = l;
}

struct RecurDev {
    const float* b_l; const uint8_t* img;
    int64_t B; int nrb;
    float* xh; float* hh; const float* c0; float* c_last;
    const float* dh; float* dz; float* dxh;
    float4* gates; float* cst; float* part;
};

// ---------------------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(RC_THREADS, 1) k_lstm_recur_fwd(const RecurDev a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x & (RC_SPLIT - 1), rb = blockIdx.x >> 3;
    const int64_t B = a.B, row0 = (int64_t)rb * 128;
    if (warp == 0) tmem_alloc<128>(&S.tmem_base);
    if (tid == 0) {
        mbar_init(&S.bar_w, 1); mbar_init(&S.bar_buf[0], 1); mbar_init(&S.bar_buf[1], 1); mbar_init(&S.bar_done, 1);
        fence_mbar_init();
        mbar_expect_tx(&S.bar_w, IMG_SLICE);
        bulk_g2s(S.W, a.img + (size_t)j * IMG_SLICE, IMG_HALF, &S.bar_w);
        bulk_g2s(S.W + IMG_HALF, a.img + (size_t)j * IMG_SLICE + IMG_HALF, IMG_HALF, &S.bar_w);
    }
    if (tid < RC_NP) S.bias[tid] = tid < RC_NL ? __ldg(a.b_l + gate_col(j, tid)) : 0.f;
    // epilogue role: TMEM lane quadrant lq (rows 32 lq ..), unit part p (7, 6, 6, 6 units)
    const int lq = warp & 3, part = warp >> 2, row = lq * 32 + lane;
    const int u_beg = part == 0 ? 0 : 7 + 6 * (part - 1), nu = part == 0 ? 7 : 6;
    const int64_t grow = row0 + row;
    const bool rvalid = grow < B;
    // loader role: two 8-float chunks per k-quarter: rows 8 rg + jj, k-chunk kg (quarter-warps take consecutive rows of one chunk
    // column => conflict-free 128-bit shared stores; the four quarter-warps take the four 32-byte pieces of the same global lines)
    int l_row[2], l_kg[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + i * RC_THREADS, blk = c >> 5;
        l_row[i] = (blk >> 1) * 8 + (c & 7);
        l_kg[i] = (blk & 1) * 4 + ((c >> 3) & 3);
    }
    float c[RC_UPT];
#pragma unroll
    for (int i = 0; i < RC_UPT; ++i) {
        c[i] = (rvalid && i < nu) ? __ldg(a.c0 + grow * RU + RC_UN * j + u_beg + i) : 0.f;
        a.cst[pidx(0, a.nrb, rb, j, i, tid)] = c[i];
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    mbar_wait(&S.bar_w, 0);
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, RC_NP);
    const uint32_t w_hi = smem_u32(S.W), w_lo = w_hi + IMG_HALF;
    uint32_t it = 0, buf_phase = 0, done_phase = 0;
    float4 ld[2][2];
    auto load_q = [&](const float* xh_t, int q) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t gr = row0 + l_row[i];
            if (gr < B) {
                const float4* p = reinterpret_cast<const float4*>(xh_t + gr * RLD + 64 * q + 8 * l_kg[i]);
                ld[i][0] = __ldcg(p); ld[i][1] = __ldcg(p + 1);
            } else {
                ld[i][0] = make_float4(0.f, 0.f, 0.f, 0.f); ld[i][1] = ld[i][0];
            }
        }
    };
#pragma unroll 1
    for (int t = 0; t < RT; ++t) {
        const float* xh_t = a.xh + (size_t)t * B * RLD;
        load_q(xh_t, 0);
#pragma unroll 1
        for (int q = 0; q < 4; ++q, ++it) {
            const uint32_t buf = it & 1u;
            if (it >= 2) {                                   // the MMAs that read this buffer two quarters ago are done
                mbar_wait(&S.bar_buf[buf], (buf_phase >> buf) & 1u);
                buf_phase ^= 1u << buf;
            }
            uint8_t* a_hi = S.A[buf];
            uint8_t* a_lo = a_hi + 16384;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint32_t h[4], l[4];
                split_pair(ld[i][0].x, ld[i][0].y, h[0], l[0]); split_pair(ld[i][0].z, ld[i][0].w, h[1], l[1]);
                split_pair(ld[i][1].x, ld[i][1].y, h[2], l[2]); split_pair(ld[i][1].z, ld[i][1].w, h[3], l[3]);
                const uint32_t off = l_kg[i] * 2048 + l_row[i] * 16;
                *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            if (q < 3) load_q(xh_t, q + 1);                  // next quarter's global loads fly over the barrier and the MMA issue
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0 && elect_one_sync()) {
                fence_after_sync();
                const uint32_t ah = smem_u32(a_hi), al = ah + 16384;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dah = make_smem_desc(ah + ks * 4096, 2048, 128), dal = make_smem_desc(al + ks * 4096, 2048, 128);
                    const uint32_t bo = (uint32_t)(8 * q + 2 * ks) * (RC_NP * 16);
                    const uint64_t dbh = make_smem_desc(w_hi + bo, RC_NP * 16, 128), dbl = make_smem_desc(w_lo + bo, RC_NP * 16, 128);
                    mma_bf16(tmem, dah, dbh, idesc, (q > 0 || ks > 0) ? 1u : 0u);
                    mma_bf16(tmem, dal, dbh, idesc, 1);
                    mma_bf16(tmem, dah, dbl, idesc, 1);
                }
                mma_commit(&S.bar_buf[buf]);
                if (q == 3) mma_commit(&S.bar_done);
            }
        }
        mbar_wait(&S.bar_done, done_phase);
        done_phase ^= 1u;
        fence_after_sync();
        // ---- cell: gates, c, m for this thread's row and units -----------------------------------------------------------------
        {
            float z[4 * RC_UPT];
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + 4u * (uint32_t)u_beg;
#pragma unroll
            for (int i = 0; i < RC_UPT; ++i)
                if (i < nu) tmem_ld_x4(taddr + 4 * i, z + 4 * i);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < RC_UPT; ++i) {
                if (i < nu) {
                    const float* bz = S.bias + 4 * (u_beg + i);
                    const float gi = sigmoid_mufu(z[4 * i] + bz[0]), gj = tanh_mufu(z[4 * i + 1] + bz[1]);
                    const float gf = sigmoid_mufu(z[4 * i + 2] + bz[2] + 1.0f), go = sigmoid_mufu(z[4 * i + 3] + bz[3]);
                    c[i] = fmaf(gf, c[i], gi * gj);
                    const float m = go * tanh_mufu(c[i]);
                    S.mst[row * RC_UN + u_beg + i] = m;
                    a.gates[pidx(t, a.nrb, rb, j, i, tid)] = make_float4(gi, gj, gf, go);
                    a.cst[pidx(t + 1, a.nrb, rb, j, i, tid)] = c[i];
                }
            }
        }
        fence_before_sync();          // this step's TMEM reads are ordered before the next step's first MMA (issued after a CTA barrier)
        __syncthreads();
        for (int idx = tid; idx < 128 * RC_UN; idx += RC_THREADS) {
            const int r = idx / RC_UN, cc = idx - r * RC_UN;
            const int64_t gr = row0 + r;
            if (gr < B) {
                const float m = S.mst[idx];
                a.hh[((size_t)t * B + gr) * RU + RC_UN * j + cc] = m;
                if (t + 1 < RT) a.xh[((size_t)(t + 1) * B + gr) * RLD + RX + RC_UN * j + cc] = m;
            }
        }
        if (t + 1 < RT) {
            cluster_sync_all();       // m_t of all 8 unit slices is in global memory before anyone loads xh_{t+1}
        } else if (a.c_last) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RC_UPT; ++i)
                if (i < nu) S.mst[row * RC_UN + u_beg + i] = c[i];
            __syncthreads();
            for (int idx = tid; idx < 128 * RC_UN; idx += RC_THREADS) {
                const int r = idx / RC_UN, cc = idx - r * RC_UN;
                const int64_t gr = row0 + r;
                if (gr < B) a.c_last[gr * RU + RC_UN * j + cc] = S.mst[idx];
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// --------------------------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(RC_THREADS, 1) k_lstm_recur_bwd(const RecurDev a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x & (RC_SPLIT - 1), rb = blockIdx.x >> 3, nrb = a.nrb;
    const int64_t B = a.B, row0 = (int64_t)rb * 128;
    if (warp == 0) tmem_alloc<256>(&S.tmem_base);
    if (tid == 0) {
        mbar_init(&S.bar_w, 1); mbar_init(&S.bar_done, 1);
        fence_mbar_init();
        mbar_expect_tx(&S.bar_w, IMG_SLICE);
        bulk_g2s(S.Bm, a.img + (size_t)j * IMG_SLICE, IMG_HALF, &S.bar_w);
        bulk_g2s(S.Bm + IMG_HALF, a.img + (size_t)j * IMG_SLICE + IMG_HALF, IMG_HALF, &S.bar_w);
    }
    for (int i = tid; i < (int)(sizeof(S.A) / 16); i += RC_THREADS) reinterpret_cast<uint4*>(S.A)[i] = make_uint4(0u, 0u, 0u, 0u);   // k = 100..111 stay zero
    const int lq = warp & 3, part = warp >> 2, row = lq * 32 + lane;
    const int u_beg = part == 0 ? 0 : 7 + 6 * (part - 1), nu = part == 0 ? 7 : 6;
    const int64_t grow = row0 + row;
    float dc[RC_UPT];
#pragma unroll
    for (int i = 0; i < RC_UPT; ++i) dc[i] = 0.f;
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    mbar_wait(&S.bar_w, 0);
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, 256);
    const uint32_t a_hi = smem_u32(S.A), a_lo = a_hi + 28672, b_hi = smem_u32(S.Bm), b_lo = b_hi + IMG_HALF;
    uint32_t done_phase = 0;
    const size_t part_tile = (size_t)RLD * 128;                                  // one CTA's partial tile: [256 columns][128 rows]
    auto part_base = [&](int buf, int jp) { return a.part + (((size_t)buf * nrb + rb) * RC_SPLIT + jp) * part_tile; };
    // dL/d(prev_pdflat embedding) of step ts = columns 11..42 of the summed partial tiles; CTA j reduces 4 of the 32 columns
    auto reduce_xpart = [&](int ts) {
        const int cc = tid >> 7, r = tid & 127, col = 11 + 4 * j + cc;
        float s = 0.f;
#pragma unroll
        for (int jp = 0; jp < RC_SPLIT; ++jp) s += __ldcg(part_base(ts & 1, jp) + (size_t)col * 128 + r);
        if (row0 + r < B) a.dxh[((size_t)ts * B + row0 + r) * RLD + col] = s;
    };
#pragma unroll 1
    for (int t = RT - 1; t >= 0; --t) {
        // (a) dL/dm_t from the heads: this CTA's 25 columns of the row block, coalesced into shared memory
        for (int idx = tid; idx < 128 * RC_UN; idx += RC_THREADS) {
            const int r = idx / RC_UN, cc = idx - r * RC_UN;
            const int64_t gr = row0 + r;
            S.S[idx] = gr < B ? __ldg(a.dh + ((size_t)t * B + gr) * RU + RC_UN * j + cc) : 0.f;
        }
        // (b) recurrent part: the 8 partial tiles of step t + 1, added in CTA order
        float rec[RC_UPT];
#pragma unroll
        for (int i = 0; i < RC_UPT; ++i) rec[i] = 0.f;
        if (t + 1 < RT) {
#pragma unroll
            for (int jp = 0; jp < RC_SPLIT; ++jp) {
                const float* pb = part_base((t + 1) & 1, jp) + (size_t)(RX + RC_UN * j + u_beg) * 128 + row;
#pragma unroll
                for (int i = 0; i < RC_UPT; ++i)
                    if (i < nu) rec[i] += __ldcg(pb + (size_t)i * 128);
            }
            reduce_xpart(t + 1);
        }
        __syncthreads();
        // (c) dz of this thread's row and units -> A tile (bf16 hi / lo), registers for the staging tile
        float dzv[4 * RC_UPT];
#pragma unroll
        for (int i = 0; i < RC_UPT; ++i) {
            if (i < nu) {
                const float4 g = a.gates[pidx(t, nrb, rb, j, i, tid)];
                const float c_prev = a.cst[pidx(t, nrb, rb, j, i, tid)], c_t = a.cst[pidx(t + 1, nrb, rb, j, i, tid)];
                const float dm = S.S[row * RC_UN + u_beg + i] + rec[i];
                const float tc = tanh_mufu(c_t);
                const float dct = fmaf(dm * g.w, fmaf(-tc, tc, 1.f), dc[i]);
                dzv[4 * i] = dct * g.y * g.x * (1.f - g.x);
                dzv[4 * i + 1] = dct * g.x * fmaf(-g.y, g.y, 1.f);
                dzv[4 * i + 2] = dct * c_prev * g.z * (1.f - g.z);
                dzv[4 * i + 3] = dm * tc * g.w * (1.f - g.w);
                dc[i] = dct * g.z;
                uint32_t h0, l0, h1, l1;
                split_pair(dzv[4 * i], dzv[4 * i + 1], h0, l0);
                split_pair(dzv[4 * i + 2], dzv[4 * i + 3], h1, l1);
                const int uu = u_beg + i;
                const uint32_t off = (uint32_t)(uu >> 1) * 2048u + (uint32_t)row * 16u + (uint32_t)(uu & 1) * 8u;
                *reinterpret_cast<uint2*>(S.A + off) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(S.A + 28672 + off) = make_uint2(l0, l1);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();                  // A tile complete; every thread has read its dL/dm values from S
        if (warp == 0 && elect_one_sync()) {
            fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < RC_NP / 16; ++ks) {
                const uint64_t dah = make_smem_desc(a_hi + ks * 4096, 2048, 128), dal = make_smem_desc(a_lo + ks * 4096, 2048, 128);
                const uint64_t dbh = make_smem_desc(b_hi + ks * 8192, 4096, 128), dbl = make_smem_desc(b_lo + ks * 8192, 4096, 128);
                mma_bf16(tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                mma_bf16(tmem, dal, dbh, idesc, 1);
                mma_bf16(tmem, dah, dbl, idesc, 1);
            }
            mma_commit(&S.bar_done);
        }
        // (d) dz in the [T*B][800] layout the weight-gradient GEMM reads, through a padded staging tile (under the MMAs)
#pragma unroll
        for (int i = 0; i < RC_UPT; ++i)
            if (i < nu) {
#pragma unroll
                for (int g = 0; g < 4; ++g) S.S[row * DZ_LD + g * RC_UN + u_beg + i] = dzv[4 * i + g];
            }
        __syncthreads();
        for (int idx = tid; idx < 128 * RC_NL; idx += RC_THREADS) {
            const int r = idx / RC_NL, rem = idx - r * RC_NL, g = rem / RC_UN, uu = rem - g * RC_UN;
            const int64_t gr = row0 + r;
            if (gr < B) a.dz[((size_t)t * B + gr) * RG + g * RU + RC_UN * j + uu] = S.S[r * DZ_LD + rem];
        }
        // (e) partial d[x | m_prev] tile of this CTA's K-slice -> global, column-major (lanes = consecutive rows)
        mbar_wait(&S.bar_done, done_phase);
        done_phase ^= 1u;
        fence_after_sync();
        {
            float* pt = part_base(t & 1, j) + row;
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + 64u * (uint32_t)part;
            float va[16], vb[16];
            tmem_ld_x16(taddr, va);
#pragma unroll
            for (int cc = 0; cc < 4; cc += 2) {
                tmem_ld_wait();
                tmem_ld_x16(taddr + 16 * (cc + 1), vb);
#pragma unroll
                for (int k = 0; k < 16; ++k) pt[(size_t)(64 * part + 16 * cc + k) * 128] = va[k];
                tmem_ld_wait();
                if (cc + 2 < 4) tmem_ld_x16(taddr + 16 * (cc + 2), va);
#pragma unroll
                for (int k = 0; k < 16; ++k) pt[(size_t)(64 * part + 16 * (cc + 1) + k) * 128] = vb[k];
            }
        }
        fence_before_sync();
        cluster_sync_all();               // partial tiles of step t are in global memory; S / A / TMEM are free again
    }
    reduce_xpart(0);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

struct Carve { uint8_t *img_f, *img_b; float4* gates; float* cst; float* part; };
inline int64_t row_blocks(int64_t B) { return (B + 127) / 128; }
inline size_t gates_floats(int64_t nrb) { return (size_t)RT * nrb * RC_SPLIT * RC_UPT * RC_THREADS * 4; }
inline size_t cst_floats(int64_t nrb) { return (size_t)(RT + 1) * nrb * RC_SPLIT * RC_UPT * RC_THREADS; }
inline size_t part_floats(int64_t nrb) { return (size_t)2 * nrb * RC_SPLIT * RLD * 128; }
constexpr size_t IMG_FLOATS = (size_t)RC_SPLIT * IMG_SLICE / 4;
Carve carve(float* s, int64_t B) {
    const int64_t nrb = row_blocks(B);
    Carve c;
    c.img_f = reinterpret_cast<uint8_t*>(s); s += IMG_FLOATS;
    c.img_b = reinterpret_cast<uint8_t*>(s); s += IMG_FLOATS;
    c.gates = reinterpret_cast<float4*>(s); s += gates_floats(nrb);
    c.cst = s; s += cst_floats(nrb);
    c.part = s;
    return c;
}
RecurDev make_dev(const LstmRecurArgs& a, bool backward) {
    const Carve c = carve(a.scratch, a.B);
    RecurDev d{};
    d.b_l = a.b_l; d.img = backward ? c.img_b : c.img_f; d.B = a.B; d.nrb = (int)row_blocks(a.B);
    d.xh = a.xh; d.hh = a.hh; d.c0 = a.c0; d.c_last = a.c_last; d.dh = a.dh; d.dz = a.dz; d.dxh = a.dxh;
    d.gates = c.gates; d.cst = c.cst; d.part = c.part;
    return d;
}
template <typename K> int launch_cluster(K kern, size_t smem, const RecurDev& d, cudaStream_t st) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(RC_SPLIT * d.nrb));
    cfg.blockDim = dim3(RC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = RC_SPLIT; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RB_CUDA(cudaLaunchKernelEx(&cfg, kern, d));
    return RB_OK;
}

}  // namespace

size_t lstm_recur_ws_floats(int64_t B) {
    const int64_t nrb = row_blocks(B);
    return 2 * IMG_FLOATS + gates_floats(nrb) + cst_floats(nrb) + part_floats(nrb) + 64;
}

int lstm_recur_build_images(const LstmRecurArgs& a, cudaStream_t st) {
    const Carve c = carve(a.scratch, a.B);
    const int total = RC_SPLIT * RLD * RC_NP;
    k_lstm_recur_images<<<(total + 255) / 256, 256, 0, st>>>(a.W_l, c.img_f, c.img_b);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}
int lstm_recur_forward(const LstmRecurArgs& a, cudaStream_t st) { return launch_cluster(k_lstm_recur_fwd, sizeof(FwdSmem), make_dev(a, false), st); }
int lstm_recur_backward(const LstmRecurArgs& a, cudaStream_t st) { return launch_cluster(k_lstm_recur_bwd, sizeof(BwdSmem), make_dev(a, true), st); }

}  // namespace rb
