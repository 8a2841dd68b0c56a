// LSTM student recurrence (/root/reference src/distilation/student_nn.py:23,41: tf.contrib.rnn.LSTMCell(200), forget_bias 1, unrolled
// over the T = 10 steps of a window; trained by back-propagation through time, lstm_train.py:74-79) as TWO persistent launches: one for
// the forward recurrence, one for BPTT.  They replace 20 dependent GEMM launches + 20 element-wise cell launches + split-K reductions.
//
// Decomposition.  Window rows are independent, time steps are not, so a *row block* of 128 windows (one M = 128 tcgen05 accumulator)
// runs all T steps inside one GROUP of 8 co-resident CTAs without any grid-wide synchronisation.  The 200 LSTM units are dealt to
// the 8 CTAs (25 each); CTA j keeps, resident in shared memory for the whole launch, the bf16 hi/lo image of ITS slice of W_l:
//   forward : z[128 x 100] = [x_t | m_{t-1}] [128 x 256] * W_l[:, cols of its 25 units x 4 gates]          (N = 112, K = 256)
//             -> gates, c (registers, carried over the steps), m -> global (hh_t for the heads, xh_{t+1} for the next step)
//   backward: dz of its 25 units from dm, dc (registers) and the saved gates -> A tile [128 x 112]; the recurrent / input gradient
//             d[x | m_prev] = dz W_l^T is a sum over ALL units, so every CTA multiplies its K-slice, dz_j [128 x 112] * W_l^T[112 x 256],
//             into a PARTIAL [128 x 256] tile (N = 256), written column-major to global; at the next step the owner of a unit adds
//             the 8 partials in CTA order (deterministic).
// The only exchange between the CTAs of a group is m_t (forward) / the partial tiles (backward), through global memory (L2), ordered
// by one group barrier per step (release / acquire on a global counter; the launch is cooperative so all CTAs are co-resident).  A first
// version used 8-CTA thread-block clusters + barrier.cluster: the device holds only 15 such clusters at a time (GPC granularity), so the
// 16 row blocks of 2048 windows took two waves; 18 groups of 8 plain CTAs fit in one (144 of 148 SMs).  Operands follow the bf16x3 scheme of the other tensor-core kernels
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
constexpr uint32_t FA_LBO = 2048 + 32, FA_HALF = 8 * FA_LBO;          // forward A tile: padded k-chunk stride (conflict-free stores, see loader)
constexpr int DZ_LD = 101;                                              // padded row of the dz staging tile (conflict-free)

struct __align__(128) FwdSmem {
    uint8_t W[IMG_SLICE];            // hi | lo : [32 k-chunks][112 rows (n = 4 u + gate)][16 B]
    uint8_t A[2][2 * FA_HALF];       // two k-quarter buffers, each hi | lo: [8 k-chunks][128 rows][16 B], chunk stride padded by 32 B
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

// barrier over the 8 CTAs of a group: every thread's global writes before it are visible to every thread of the group after it
__device__ __forceinline__ void group_barrier(unsigned int* cnt, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory"); } while (v < target);
    }
    __syncthreads();
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

// phase timestamps (globaltimer, ns) of CTA 0 of the last launches: [0] forward, [1] backward; 16 slots per step
__device__ unsigned long long g_recur_stamps[2][16 * RT];
__device__ __forceinline__ void rc_stamp(int k, int t, int i) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        g_recur_stamps[k][16 * t + i] = v;
    }
}

struct RecurDev {
    const float* b_l; const uint8_t* img;
    int64_t B; int nrb;
    float* xh; float* hh; int hh_ld; const float* c0; float* c_last;
    const float* dh; float* dz; float* dxh;
    float4* gates; float* cst; float* part;
    unsigned int* bar;       // one counter per group (zeroed by lstm_recur_build_images)
    int ngroups;
};

// ---------------------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(RC_THREADS, 1) k_lstm_recur_fwd(const RecurDev a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x & (RC_SPLIT - 1), grp = blockIdx.x >> 3;
    const int64_t B = a.B;
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
    // loader role.  A k-quarter of the A tile is 128 rows x 256 B of xh.  Global side: the 8 lanes of a quarter-warp read the 8
    // consecutive 16-byte pieces of ONE 128-byte line (the L1 handles a 128-bit load per quarter-warp: 8 different lines per quarter-warp,
    // as a row-per-lane mapping gives, cost 8 tag look-ups instead of 1 and made this phase 5 us per step), first of row r0, then of row
    // r0 + 1.  Neighbouring lanes then swap one piece so that even lanes own the 32-byte chunk (r0, kg) and odd lanes (r0 + 1, kg).
    // Shared side: chunk (row, kg) at kg * FA_LBO + row * 16; the 32-byte pad makes a quarter-warp's 2 rows x 4 chunks hit 8 distinct
    // 16-byte bank groups.
    const int l8 = lane & 7, qw = lane >> 3;
    int l_r0[2], l_piece, l_row[2], l_kg;
    l_piece = 8 * (qw & 1) + l8;                           // 16-byte piece of the row's 256-byte quarter
    l_kg = 4 * (qw & 1) + (l8 >> 1);                       // chunk this lane owns after the swap
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        l_r0[i] = 8 * warp + 4 * i + 2 * (qw >> 1);
        l_row[i] = l_r0[i] + (l8 & 1);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    mbar_wait(&S.bar_w, 0);
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, RC_NP);
    const uint32_t w_hi = smem_u32(S.W), w_lo = w_hi + IMG_HALF;
    unsigned int* bar = a.bar + grp;
    uint32_t it = 0, buf_phase = 0, done_phase = 0, epoch = 0;
#pragma unroll 1
    for (int rb = grp; rb < a.nrb; rb += a.ngroups) {
        const int64_t row0 = (int64_t)rb * 128, grow = row0 + row;
        const bool rvalid = grow < B;
        float c[RC_UPT];
#pragma unroll
        for (int i = 0; i < RC_UPT; ++i) {
            c[i] = (rvalid && i < nu) ? __ldg(a.c0 + grow * RU + RC_UN * j + u_beg + i) : 0.f;
            a.cst[pidx(0, a.nrb, rb, j, i, tid)] = c[i];
        }
        float4 ld[3][2][2];          // three k-quarters of global loads in flight (registers): [slot][row pair][row r0 / r0 + 1]
        auto load_q = [&](const float* xh_t, int q, float4 (&dst)[2][2]) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t gr = row0 + l_r0[i] + h;
                    dst[i][h] = gr < B ? __ldcg(reinterpret_cast<const float4*>(xh_t + gr * RLD + 64 * q) + l_piece) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
#pragma unroll 1
        for (int t = 0; t < RT; ++t) {
            const float* xh_t = a.xh + (size_t)t * B * RLD;
            rc_stamp(0, t, 0);
            load_q(xh_t, 0, ld[0]); load_q(xh_t, 1, ld[1]); load_q(xh_t, 2, ld[2]);
#pragma unroll
            for (int q = 0; q < 4; ++q, ++it) {
                const uint32_t buf = it & 1u;
                if (it >= 2) {                               // the MMAs that read this buffer two quarters ago are done
                    mbar_wait(&S.bar_buf[buf], (buf_phase >> buf) & 1u);
                    buf_phase ^= 1u << buf;
                }
                uint8_t* a_hi = S.A[buf];
                uint8_t* a_lo = a_hi + FA_HALF;
                float4 (&cur)[2][2] = ld[q % 3];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    // even lane: keeps its piece of row r0, gets the neighbour's piece of row r0 (second half of the chunk);
                    // odd lane : keeps its piece of row r0 + 1, gets the neighbour's piece of row r0 + 1 (first half of the chunk)
                    const bool odd = (l8 & 1) != 0;
                    const float4 give = odd ? cur[i][0] : cur[i][1], keep = odd ? cur[i][1] : cur[i][0];
                    float4 got;
                    got.x = __shfl_xor_sync(0xffffffffu, give.x, 1); got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
                    got.z = __shfl_xor_sync(0xffffffffu, give.z, 1); got.w = __shfl_xor_sync(0xffffffffu, give.w, 1);
                    const float4 lo4 = odd ? got : keep, hi4 = odd ? keep : got;       // first / second 16 bytes of the chunk
                    uint32_t h[4], l[4];
                    split_pair(lo4.x, lo4.y, h[0], l[0]); split_pair(lo4.z, lo4.w, h[1], l[1]);
                    split_pair(hi4.x, hi4.y, h[2], l[2]); split_pair(hi4.z, hi4.w, h[3], l[3]);
                    const uint32_t off = l_kg * FA_LBO + l_row[i] * 16;
                    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                }
                if (q == 0) load_q(xh_t, 3, ld[0]);          // the last quarter's loads fly over the first three quarters
                fence_async_smem();
                fence_before_sync();
                __syncthreads();
                if (warp == 0 && elect_one_sync()) {
                    fence_after_sync();
                    const uint32_t ah = smem_u32(a_hi), al = ah + FA_HALF;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dah = make_smem_desc(ah + ks * 2 * FA_LBO, FA_LBO, 128), dal = make_smem_desc(al + ks * 2 * FA_LBO, FA_LBO, 128);
                        const uint32_t bo = (uint32_t)(8 * q + 2 * ks) * (RC_NP * 16);
                        const uint64_t dbh = make_smem_desc(w_hi + bo, RC_NP * 16, 128), dbl = make_smem_desc(w_lo + bo, RC_NP * 16, 128);
                        mma_bf16(tmem, dah, dbh, idesc, (q > 0 || ks > 0) ? 1u : 0u);
                        mma_bf16(tmem, dal, dbh, idesc, 1);
                        mma_bf16(tmem, dah, dbl, idesc, 1);
                    }
                    mma_commit(&S.bar_buf[buf]);
                    if (q == 3) mma_commit(&S.bar_done);
                }
                rc_stamp(0, t, 1 + q);
            }
            mbar_wait(&S.bar_done, done_phase);
            done_phase ^= 1u;
            fence_after_sync();
            rc_stamp(0, t, 5);
            // ---- cell: gates, c, m for this thread's row and units -------------------------------------------------------------
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
            fence_before_sync();      // this step's TMEM reads are ordered before the next step's first MMA (issued after a CTA barrier)
            __syncthreads();
            rc_stamp(0, t, 6);
            if (lane < RC_UN) {       // warp w stores rows w, w + 16, ...: one 100-byte segment per row and destination
                float* hh_t = a.hh + ((size_t)t * B + row0) * a.hh_ld + RC_UN * j + lane;
                float* xh_n = a.xh + ((size_t)(t + 1) * B + row0) * RLD + RX + RC_UN * j + lane;
#pragma unroll
                for (int r = warp; r < 128; r += RC_THREADS / 32) {
                    if (row0 + r < B) {
                        const float m = S.mst[r * RC_UN + lane];
                        hh_t[(size_t)r * a.hh_ld] = m;
                        if (t + 1 < RT) xh_n[(size_t)r * RLD] = m;
                    }
                }
            }
            rc_stamp(0, t, 7);
            if (t + 1 < RT) {
                group_barrier(bar, RC_SPLIT * (++epoch));       // m_t of all 8 unit slices is in global memory before anyone loads xh_{t+1}
                rc_stamp(0, t, 8);
            } else {
                __syncthreads();
                if (a.c_last) {
#pragma unroll
                    for (int i = 0; i < RC_UPT; ++i)
                        if (i < nu) S.mst[row * RC_UN + u_beg + i] = c[i];
                    __syncthreads();
                    if (lane < RC_UN) {
#pragma unroll
                        for (int r = warp; r < 128; r += RC_THREADS / 32)
                            if (row0 + r < B) a.c_last[(row0 + r) * RU + RC_UN * j + lane] = S.mst[r * RC_UN + lane];
                    }
                    __syncthreads();
                }
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
    const int j = blockIdx.x & (RC_SPLIT - 1), grp = blockIdx.x >> 3, nrb = a.nrb;
    const int64_t B = a.B;
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
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    mbar_wait(&S.bar_w, 0);
    const uint32_t tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, 256);
    const uint32_t a_hi = smem_u32(S.A), a_lo = a_hi + 28672, b_hi = smem_u32(S.Bm), b_lo = b_hi + IMG_HALF;
    unsigned int* bar = a.bar + a.ngroups + grp;
    uint32_t done_phase = 0, epoch = 0;
    const size_t part_tile = (size_t)RLD * 128;                                  // one CTA's partial tile: [256 columns][128 rows]
#pragma unroll 1
    for (int rb = grp; rb < nrb; rb += a.ngroups) {
        const int64_t row0 = (int64_t)rb * 128;
        auto part_base = [&](int buf, int jp) { return a.part + (((size_t)buf * nrb + rb) * RC_SPLIT + jp) * part_tile; };
        // dL/d(prev_pdflat embedding) of step ts = columns 11..42 of the summed partial tiles; CTA j reduces 4 of the 32 columns
        auto reduce_xpart = [&](int ts) {
            const int cc = tid >> 7, r = tid & 127, col = 11 + 4 * j + cc;
            float s = 0.f;
#pragma unroll
            for (int jp = 0; jp < RC_SPLIT; ++jp) s += __ldcg(part_base(ts & 1, jp) + (size_t)col * 128 + r);
            if (row0 + r < B) a.dxh[((size_t)ts * B + row0 + r) * RLD + col] = s;
        };
        float dc[RC_UPT];
#pragma unroll
        for (int i = 0; i < RC_UPT; ++i) dc[i] = 0.f;
#pragma unroll 1
        for (int t = RT - 1; t >= 0; --t) {
            rc_stamp(1, t, 0);
            // (a) loads that do not depend on the recurrence go first: saved gates / cell states, dL/dm_t from the heads
            float4 g[RC_UPT];
            float c_prev[RC_UPT], c_t[RC_UPT];
#pragma unroll
            for (int i = 0; i < RC_UPT; ++i) {
                if (i < nu) {
                    g[i] = __ldcs(a.gates + pidx(t, nrb, rb, j, i, tid));
                    c_prev[i] = __ldcs(a.cst + pidx(t, nrb, rb, j, i, tid));
                    c_t[i] = __ldcs(a.cst + pidx(t + 1, nrb, rb, j, i, tid));
                }
            }
            if (lane < RC_UN) {
                const float* dh_t = a.dh + ((size_t)t * B + row0) * RU + RC_UN * j + lane;
#pragma unroll
                for (int r = warp; r < 128; r += RC_THREADS / 32) S.S[r * RC_UN + lane] = row0 + r < B ? __ldg(dh_t + (size_t)r * RU) : 0.f;
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
            rc_stamp(1, t, 1);
            // (c) dz of this thread's row and units -> A tile (bf16 hi / lo), registers for the staging tile
            float dzv[4 * RC_UPT];
#pragma unroll
            for (int i = 0; i < RC_UPT; ++i) {
                if (i < nu) {
                    const float dm = S.S[row * RC_UN + u_beg + i] + rec[i];
                    const float tc = tanh_mufu(c_t[i]);
                    const float dct = fmaf(dm * g[i].w, fmaf(-tc, tc, 1.f), dc[i]);
                    dzv[4 * i] = dct * g[i].y * g[i].x * (1.f - g[i].x);
                    dzv[4 * i + 1] = dct * g[i].x * fmaf(-g[i].y, g[i].y, 1.f);
                    dzv[4 * i + 2] = dct * c_prev[i] * g[i].z * (1.f - g[i].z);
                    dzv[4 * i + 3] = dm * tc * g[i].w * (1.f - g[i].w);
                    dc[i] = dct * g[i].z;
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
            __syncthreads();              // A tile complete; every thread has read its dL/dm values from S
            rc_stamp(1, t, 2);
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
                    for (int gg = 0; gg < 4; ++gg) S.S[row * DZ_LD + gg * RC_UN + u_beg + i] = dzv[4 * i + gg];
                }
            __syncthreads();
            rc_stamp(1, t, 3);
            {                             // warp w stores rows w, w + 16, ...: four 100-byte segments (gates) per row
                float* dz_t = a.dz + ((size_t)t * B + row0) * RG + RC_UN * j;
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const int rem = lane + 32 * pass, gg = rem / RC_UN, coff = gg * RU + (rem - gg * RC_UN);
                    if (rem < RC_NL) {
#pragma unroll
                        for (int r = warp; r < 128; r += RC_THREADS / 32)
                            if (row0 + r < B) dz_t[(size_t)r * RG + coff] = S.S[r * DZ_LD + rem];
                    }
                }
            }
            // (e) partial d[x | m_prev] tile of this CTA's K-slice -> global, column-major (lanes = consecutive rows)
            rc_stamp(1, t, 4);
            mbar_wait(&S.bar_done, done_phase);
            done_phase ^= 1u;
            fence_after_sync();
            rc_stamp(1, t, 5);
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
            rc_stamp(1, t, 6);
            group_barrier(bar, RC_SPLIT * (++epoch));     // partial tiles of step t are in global memory; S / A / TMEM are free again
            rc_stamp(1, t, 7);
        }
        reduce_xpart(0);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

constexpr int RC_MAX_GROUPS = 64;
struct Carve { uint8_t *img_f, *img_b; float4* gates; float* cst; float* part; unsigned int* bar; };
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
    c.part = s; s += part_floats(nrb);
    c.bar = reinterpret_cast<unsigned int*>(s);
    return c;
}
// groups of 8 CTAs that are co-resident (one CTA per SM: the weight slice fills the shared memory)
int group_count(int64_t nrb, int* out) {
    int device = 0, sms = 0;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int64_t g = sms / RC_SPLIT;
    RB_REQUIRE(g >= 1, "the LSTM recurrence kernels need at least 8 SMs");
    *out = (int)(nrb < g ? nrb : (g < RC_MAX_GROUPS ? g : RC_MAX_GROUPS));
    return RB_OK;
}
int make_dev(const LstmRecurArgs& a, bool backward, RecurDev& d) {
    const Carve c = carve(a.scratch, a.B);
    d = RecurDev{};
    d.b_l = a.b_l; d.img = backward ? c.img_b : c.img_f; d.B = a.B; d.nrb = (int)row_blocks(a.B);
    d.xh = a.xh; d.hh = a.hh; d.hh_ld = a.hh_ld; d.c0 = a.c0; d.c_last = a.c_last; d.dh = a.dh; d.dz = a.dz; d.dxh = a.dxh;
    d.gates = c.gates; d.cst = c.cst; d.part = c.part; d.bar = c.bar;
    return group_count(d.nrb, &d.ngroups);
}
template <typename K> int launch_groups(K kern, size_t smem, RecurDev& d, cudaStream_t st) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&d};
    // cooperative: the per-step group barriers spin on a global counter, so all CTAs must be co-resident
    RB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(RC_SPLIT * d.ngroups)), dim3(RC_THREADS), args, smem, st));
    return RB_OK;
}

}  // namespace

size_t lstm_recur_ws_floats(int64_t B) {
    const int64_t nrb = row_blocks(B);
    return 2 * IMG_FLOATS + gates_floats(nrb) + cst_floats(nrb) + part_floats(nrb) + 2 * RC_MAX_GROUPS + 64;
}

int lstm_recur_build_images(const LstmRecurArgs& a, cudaStream_t st) {
    const Carve c = carve(a.scratch, a.B);
    const int total = RC_SPLIT * RLD * RC_NP;
    RB_CUDA(cudaMemsetAsync(c.bar, 0, sizeof(unsigned int) * 2 * RC_MAX_GROUPS, st));      // group-barrier counters of the two launches
    k_lstm_recur_images<<<(total + 255) / 256, 256, 0, st>>>(a.W_l, c.img_f, c.img_b);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}
int lstm_recur_forward(const LstmRecurArgs& a, cudaStream_t st) {
    RecurDev d;
    int rc = make_dev(a, false, d);
    return rc ? rc : launch_groups(k_lstm_recur_fwd, sizeof(FwdSmem), d, st);
}
int lstm_recur_backward(const LstmRecurArgs& a, cudaStream_t st) {
    RecurDev d;
    int rc = make_dev(a, true, d);
    return rc ? rc : launch_groups(k_lstm_recur_bwd, sizeof(BwdSmem), d, st);
}

}  // namespace rb

// debug: phase stamps of CTA 0 of the last forward / backward recurrence launch (2 x 16 x T values, ns)
extern "C" int rb_debug_lstm_recur_stamps(unsigned long long* host_out) {
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpyFromSymbol(host_out, rb::g_recur_stamps, sizeof(unsigned long long) * 2 * 16 * rb::RT));
    return RB_OK;
}
