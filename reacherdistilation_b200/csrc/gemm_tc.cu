// General fp32-in / fp32-out GEMM on the 5th-gen tensor cores with bf16x3 operand splitting:
//     C[M,N] (+)= epilogue( A[M,K] * B[K,N] )        epilogue: + bias[n], tanh, * (1 - H[m,n]^2)
// Building block of the LSTM student (/root/reference src/distilation/student_nn.py:21-49: LSTMCell(200) and the per-step dense
// heads; forward, dgrad and wgrad are all calls of this kernel with different operand orientations).
//
// Each operand is described by where its CONTIGUOUS dimension lies in global memory:
//     K-contiguous  (x_mn = 0): element (row, k) at X[row * ld + k]   -> UMMA K-major tile
//     MN-contiguous (x_mn = 1): element (row, k) at X[k * ld + row]   -> UMMA MN-major tile (no transpose pass needed)
// The loader threads read 8 contiguous floats (one 32 B sector), split them into bf16 hi / lo and write one 16-byte chunk into the
// no-swizzle canonical layout of that orientation, so A^T B (wgrad), A B^T (dgrad) and A B (forward) all run at the same speed.
// 128 x BN output tile per CTA, BK = 32 per k-tile.  Pipeline: cp.async (LDGSTS) streams the raw fp32 k-tiles into 2-4 shared-memory
// stages; each k-tile is then split into bf16 hi / lo operand stages (1-2, reuse gated by tcgen05.commit -> mbarrier) and consumed by MMAs
// issued by one elected thread; accumulator in TMEM; the epilogue goes through a padded shared-memory tile so every global access is a
// coalesced 128-bit one.  Two pipeline shapes (GemmCfg): DEEP for grids of at most one CTA per SM, PACKED (two / three CTAs per SM) above.
// Where a 128-wide CTA of the LSTM heads (K = 64) spends its 7 us (rb_debug_gemm_stamps, B200): set-up 0.9, first copies 0.7, 2 x [convert +
// barrier] 0.4, MMA tail 0.3, TMEM -> shared 0.5, output 3.3 (10 MB of fp32 rows from 160 CTAs at once: store-bound), i.e. fixed costs, not MMAs.
// Split-K (grid.z) writes partial tiles that a second kernel adds in a fixed order (deterministic).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace rb {

using namespace tc;

constexpr int GM_THREADS = 256;
constexpr int GM_BM = 128;
constexpr int GM_BK = 32;
constexpr int GM_BF_STAGES_MAX = 2;      // bf16 hi/lo operand stages (consumed by the MMAs): see GemmCfg

struct __align__(16) GemmCtl {
    uint64_t stage_bar[GM_BF_STAGES_MAX];
    uint64_t done_bar;
    uint32_t tmem_base;
};

// ---- global -> shared, asynchronously (LDGSTS).  A chunk = 8 consecutive fp32 of the operand's contiguous dimension; its two 16-byte
// halves go to two planes of the raw stage (plane[h][chunk]: 16-byte stride, conflict-free both ways).  Elements outside the
// operand are zero-filled by the copy itself (src-size < cp-size).
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-thread descriptor of one chunk of an operand tile, set up ONCE before the k-loop: where it comes from (pointer at k-tile 0, advanced
// by a constant step per k-tile), how many of its 8 elements exist along the row dimension, and where it goes (raw plane slot, bf16 slot).
struct ChunkPlan {
    const float* p;       // source of the chunk in k-tile 0 (advanced by `step` floats per k-tile)
    int kofs;             // k of the chunk inside the k-tile
    int rows_ok;          // K-major: 1 if the row exists; MN-major: how many of the 8 rows exist (0..8)
    uint32_t raw, off;    // byte offsets: raw plane slot (c * 16), bf16 tile slot
    bool aligned, active;
};
// bf16 tile geometry (one half, hi or lo) of an operand with R rows, BK = 32, no-swizzle canonical layouts with PADDED strides:
//   K-major : chunk (row r, k-group kg)  at kg * LBO_K + r * 16,                    LBO_K = R*16 + 32, SBO = 128
//   MN-major: chunk (k, row-group rg)    at rg * SBO_MN + (k/8)*128 + (k%8)*16,     LBO = 128, SBO_MN = 512 + 16
// The pads let a quarter-warp whose 8 lanes read ONE or TWO contiguous global lines (see plan_chunk) store its 8 chunks into 8 distinct
// 16-byte bank groups.
template <int R> struct TileGeo {
    static constexpr int LBO_K = R * 16 + 32;
    static constexpr int SBO_MN = 512 + 16;
    static constexpr int BYTES_RAW = (4 * LBO_K > (R / 8) * SBO_MN) ? 4 * LBO_K : (R / 8) * SBO_MN;
    static constexpr int BYTES = (BYTES_RAW + 127) / 128 * 128;
};
template <int R>
__device__ __forceinline__ void plan_chunk(ChunkPlan& cp, int c, const float* __restrict__ X, int ld, int mn, int row0, int nrows, int kbeg, bool active) {
    // Chunk -> lane mapping.  The L1 handles a 128-bit-per-lane request one QUARTER-warp (8 lanes) at a time and pays one tag look-up per
    // distinct 128-byte line in it: a row-per-lane mapping (8 lines per quarter-warp) made the copies of a k-tile cost ~2000 cycles.  Here
    // a quarter-warp (task) takes chunks that lie in one or two contiguous lines:
    //   K-major : 2 consecutive rows x their 4 k-groups (two 128-byte lines)
    //   MN-major: 8 consecutive row-groups at one k (256 contiguous bytes); narrower tiles: R/8 row-groups x 8/(R/8) values of k
    int row, k;
    const int l8 = c & 7, task = c >> 3;
    active = active && task < R / 2;
    if (mn == 0) {
        row = 2 * task + (l8 >> 2); k = (l8 & 3) * 8;
        cp.off = (l8 & 3) * TileGeo<R>::LBO_K + row * 16;
    } else {
        constexpr int RGn = R / 8, LPK = RGn < 8 ? RGn : 8, KS = LPK;          // lanes per k; k stride inside a task = 8 / (8 / LPK)
        int rg;
        if (RGn >= 8) { k = task & 31; rg = (task >> 5) * 8 + l8; }
        else { k = (task % KS) + 8 * (task / KS) + KS * (l8 / LPK); rg = l8 % LPK; }
        row = rg * 8;
        cp.off = rg * TileGeo<R>::SBO_MN + (k >> 3) * 128 + (k & 7) * 16;
    }
    cp.kofs = k; cp.raw = c * 16; cp.active = active;
    const int gr = row0 + row;
    if (mn == 0) { cp.rows_ok = gr < nrows ? 1 : 0; cp.p = X + (size_t)min(gr, nrows - 1) * ld + kbeg + k; }
    else { cp.rows_ok = min(8, max(0, nrows - gr)); cp.p = X + (size_t)(kbeg + k) * ld + min(gr, nrows - 1); }
    cp.aligned = (reinterpret_cast<uintptr_t>(cp.p) & 15) == 0;        // the per-k-tile step (128 B or 128 * ld B) keeps it
}
// copy the chunk of k-tile starting at k0 (absolute) into its raw slots; elements at k >= kend or beyond the rows are zero-filled by the copy
__device__ __forceinline__ void copy_planned(const ChunkPlan& cp, const float* X, int mn, int k0, int kend, uint32_t plane0, uint32_t plane1) {
    int cnt;
    if (mn == 0) cnt = cp.rows_ok ? min(8, max(0, kend - (k0 + cp.kofs))) : 0;
    else cnt = (k0 + cp.kofs < kend) ? cp.rows_ok : 0;
    const uint32_t d0 = plane0 + cp.raw, d1 = plane1 + cp.raw;
    if (cnt == 0) {                                                 // nothing is read (src-size 0): any 16-byte aligned address inside the allocation
        const float* z = reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(X) & ~(uintptr_t)15);
        cp_async_16(d0, z, 0);
        cp_async_16(d1, z, 0);
        return;
    }
    const float* p = cp.p;
    if (cp.aligned) {
        cp_async_16(d0, p, min(16, 4 * cnt));
        cp_async_16(d1, cnt > 4 ? p + 4 : p, max(0, min(16, 4 * (cnt - 4))));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async_4(d0 + 4 * i, i < cnt ? p + i : p, i < cnt ? 4 : 0);
#pragma unroll
        for (int i = 4; i < 8; ++i) cp_async_4(d1 + 4 * (i - 4), i < cnt ? p + i : p, i < cnt ? 4 : 0);
    }
}
__device__ __forceinline__ void store_chunk(uint8_t* hi, uint8_t* lo, uint32_t off, const float4& a, const float4& b) {
    uint32_t h[4], l[4];
    split_pair(a.x, a.y, h[0], l[0]); split_pair(a.z, a.w, h[1], l[1]);
    split_pair(b.x, b.y, h[2], l[2]); split_pair(b.z, b.w, h[3], l[3]);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int R> struct OperandTile {
    static constexpr int CHUNKS = R * GM_BK / 8;                                 // 8-float chunks per k-tile
    static constexpr int SLOTS = CHUNKS < 128 ? 128 : CHUNKS;                    // thread slots (the MN-major mapping needs 4 warps at least)
    static constexpr int PER_THREAD = (SLOTS + GM_THREADS - 1) / GM_THREADS;
    static constexpr int BYTES = TileGeo<R>::BYTES;      // one bf16 half
    static constexpr int RAW_BYTES = SLOTS * 32;         // fp32 staging: two planes of SLOTS x 16 bytes
    __device__ static __forceinline__ uint64_t desc(uint32_t base, int mn, int ks) {       // K-step ks (16 values of k) of the tile
        return mn == 0 ? make_smem_desc(base + ks * 2 * TileGeo<R>::LBO_K, TileGeo<R>::LBO_K, 128)
                       : make_smem_desc(base + ks * 256, 128, TileGeo<R>::SBO_MN);
    }
};

// Two pipeline shapes per tile width.  DEEP: 2 bf16 stages and 2-4 raw (fp32) stages, one CTA per SM for the 128-wide tile, two for the
// narrower ones.  PACKED: 2 raw stages + 1 bf16 stage (97 / 73 / 61 KB) so that two 128-wide or three narrower CTAs share an SM.  The conversion
// of a packed CTA's next k-tile waits for the MMAs of its previous one, but the other CTAs of the SM fill that time: the per-k-tile cost is the
// latency of [copy landed -> convert -> barrier -> MMA], and 16-24 warps hide it better than 8-16; grids of 149 .. 296 wide tiles (the batched
// LSTM heads: 160) also become one wave instead of two.  Measured (B200, scripts/r02/gemm_packing_ab.py and a sweep over every {2,3,4} x {1,2}
// stage combination per width): LSTM step at 2048 windows 0.526 -> 0.487 ms, two-headed LSTM 0.620 -> 0.550 ms, 8192 x 4096 x 4096 product
// 116 -> 152 TFLOP/s; grids of at most one CTA per SM (20 / 100 windows) are insensitive (all combinations within 1 %) and keep DEEP.
template <int BN, bool PACKED = false> struct GemmCfg {
    static constexpr int RAW_STAGES = PACKED ? 2 : (BN >= 128 ? 4 : (BN == 64 ? 2 : 3));
    static constexpr int BF_STAGES = PACKED ? 1 : 2;
    static constexpr int BF_STAGE_BYTES = 2 * OperandTile<GM_BM>::BYTES + 2 * OperandTile<BN>::BYTES;
    static constexpr int RAW_STAGE_BYTES = OperandTile<GM_BM>::RAW_BYTES + OperandTile<BN>::RAW_BYTES;
    static constexpr int PIPE_BYTES = BF_STAGES * BF_STAGE_BYTES + RAW_STAGES * RAW_STAGE_BYTES;
    static constexpr int CT_LD = BN + 4;                                 // fp32 C staging tile, padded rows (conflict-free 128-bit accesses)
    static constexpr int CT_BYTES = GM_BM * CT_LD * 4;
    static constexpr int SMEM = (PIPE_BYTES > CT_BYTES ? PIPE_BYTES : CT_BYTES) + (int)sizeof(GemmCtl);
    // co-resident CTAs: what 228 KB of shared memory per SM holds (1 KB reserved per CTA), at most 3 (85 registers per thread)
    static constexpr int FIT = 233472 / (SMEM + 1024);
    static constexpr int MIN_CTAS = FIT < 1 ? 1 : (FIT > 3 ? 3 : FIT);
};
static_assert(GemmCfg<128, false>::MIN_CTAS == 1 && GemmCfg<64, false>::MIN_CTAS == 2 && GemmCfg<128, true>::MIN_CTAS == 2 &&
              GemmCfg<64, true>::MIN_CTAS == 3 && GemmCfg<32, true>::MIN_CTAS == 3 && GemmCfg<16, true>::MIN_CTAS == 3, "CTAs per SM as documented");

// phase timestamps (globaltimer, ns) of CTA 0 of the last launch: read back with rb_debug_gemm_stamps (scripts/r02/gemm_stamps.py)
__device__ unsigned long long g_gemm_stamps[16];
__device__ __forceinline__ void gm_stamp(int i) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        g_gemm_stamps[i] = v;
    }
    if ((i == 0 || i == 9) && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1 && threadIdx.x == 0) {
        unsigned long long v;                      // start / end of the LAST CTA of the grid: tells one wave from two
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        g_gemm_stamps[i == 0 ? 10 : 11] = v;
    }
}

template <int BN, bool PACKED>
__global__ void __launch_bounds__(GM_THREADS, GemmCfg<BN, PACKED>::MIN_CTAS) k_gemm_bf16x3(const GemmArgs g) {
    using TA = OperandTile<GM_BM>;
    using TB = OperandTile<BN>;
    using CF = GemmCfg<BN, PACKED>;
    constexpr int RS = CF::RAW_STAGES, BFS = CF::BF_STAGES;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* raw_base = smem + BFS * CF::BF_STAGE_BYTES;
    GemmCtl& ctl = *reinterpret_cast<GemmCtl*>(smem + CF::SMEM - sizeof(GemmCtl));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * BN;
    const int bz = blockIdx.z / g.nsplit, sz = blockIdx.z - bz * g.nsplit;          // batch index, K slice
    const float* __restrict__ Ab = g.A + (size_t)bz * g.sA;
    const float* __restrict__ Bb = g.B + (size_t)bz * g.sB;
    const int kbeg = sz * g.ksplit, kend = min(g.K, kbeg + g.ksplit);
    const int ktiles = (kend - kbeg + GM_BK - 1) / GM_BK;
    gm_stamp(0);

    if (warp == 0) tmem_alloc<TCOLS>(&ctl.tmem_base);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < BFS; ++s) mbar_init(&ctl.stage_bar[s], 1);
        mbar_init(&ctl.done_bar, 1);
        fence_mbar_init();
    }
    // every thread copies / converts the same chunks of every k-tile, so it only ever waits for its own cp.async groups
    ChunkPlan pa[TA::PER_THREAD], pb[TB::PER_THREAD];
#pragma unroll
    for (int i = 0; i < TA::PER_THREAD; ++i) plan_chunk<GM_BM>(pa[i], tid + i * GM_THREADS, Ab, g.lda, g.a_mn, m0, g.M, kbeg, tid + i * GM_THREADS < TA::SLOTS);
#pragma unroll
    for (int i = 0; i < TB::PER_THREAD; ++i) plan_chunk<BN>(pb[i], tid + i * GM_THREADS, Bb, g.ldb, g.b_mn, n0, g.N, kbeg, tid + i * GM_THREADS < TB::SLOTS);
    const size_t step_a = g.a_mn == 0 ? (size_t)GM_BK : (size_t)GM_BK * g.lda, step_b = g.b_mn == 0 ? (size_t)GM_BK : (size_t)GM_BK * g.ldb;
    auto issue_copies = [&](int kt) {                                // must be called for kt = 0, 1, 2, ... in order (the plans advance)
        const int k0 = kbeg + kt * GM_BK;
        const uint32_t ra = smem_u32(raw_base + (kt % RS) * CF::RAW_STAGE_BYTES), rb_ = ra + TA::RAW_BYTES;
#pragma unroll
        for (int i = 0; i < TA::PER_THREAD; ++i)
            if (pa[i].active) { copy_planned(pa[i], Ab, g.a_mn, k0, kend, ra, ra + TA::SLOTS * 16); pa[i].p += step_a; }
#pragma unroll
        for (int i = 0; i < TB::PER_THREAD; ++i)
            if (pb[i].active) { copy_planned(pb[i], Bb, g.b_mn, k0, kend, rb_, rb_ + TB::SLOTS * 16); pb[i].p += step_b; }
    };
#pragma unroll
    for (int p = 0; p < RS - 1; ++p) {                 // prologue: RS - 1 tiles in flight (empty groups keep the group count uniform)
        if (p < ktiles) issue_copies(p);
        cp_async_commit();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    gm_stamp(1);
    const uint32_t tmem = ctl.tmem_base;
    const uint32_t idesc = make_idesc_bf16(GM_BM, BN) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16);

    uint32_t stage_phase = 0;           // bit s = parity to wait for on stage_bar[s]
    for (int kt = 0; kt < ktiles; ++kt) {
        if (kt + RS - 1 < ktiles) issue_copies(kt + RS - 1);          // its raw stage was converted by this same thread in iteration kt - 1
        cp_async_commit();
        cp_async_wait<RS - 1>();                                      // this thread's copies of tile kt have landed
        const int s = kt % BFS;
        uint8_t* a_hi = smem + s * CF::BF_STAGE_BYTES;
        uint8_t* a_lo = a_hi + TA::BYTES;
        uint8_t* b_hi = a_lo + TA::BYTES;
        uint8_t* b_lo = b_hi + TB::BYTES;
        if (kt >= BFS) {                                              // the MMAs that read this bf16 stage (tile kt - BFS) are done
            mbar_wait(&ctl.stage_bar[s], (stage_phase >> s) & 1u);
            stage_phase ^= 1u << s;
        }
        if (kt < 2) gm_stamp(2 + 2 * kt);
        const uint8_t* ra = raw_base + (kt % RS) * CF::RAW_STAGE_BYTES;
        const uint8_t* rb_ = ra + TA::RAW_BYTES;
#pragma unroll
        for (int i = 0; i < TA::PER_THREAD; ++i)
            if (pa[i].active)
                store_chunk(a_hi, a_lo, pa[i].off, *reinterpret_cast<const float4*>(ra + pa[i].raw), *reinterpret_cast<const float4*>(ra + TA::SLOTS * 16 + pa[i].raw));
#pragma unroll
        for (int i = 0; i < TB::PER_THREAD; ++i)
            if (pb[i].active)
                store_chunk(b_hi, b_lo, pb[i].off, *reinterpret_cast<const float4*>(rb_ + pb[i].raw), *reinterpret_cast<const float4*>(rb_ + TB::SLOTS * 16 + pb[i].raw));
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (kt < 2) gm_stamp(3 + 2 * kt);
        if (warp == 0 && elect_one_sync()) {
            fence_after_sync();
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int ks = 0; ks < GM_BK / 16; ++ks) {
                const uint64_t dah = TA::desc(ah, g.a_mn, ks), dal = TA::desc(al, g.a_mn, ks), dbh = TB::desc(bh, g.b_mn, ks),
                               dbl = TB::desc(bl, g.b_mn, ks);
                mma_bf16(tmem, dah, dbh, idesc, (kt > 0 || ks > 0) ? 1u : 0u);
                mma_bf16(tmem, dal, dbh, idesc, 1);
                mma_bf16(tmem, dah, dbl, idesc, 1);
            }
            mma_commit(&ctl.stage_bar[s]);
            if (kt == ktiles - 1) mma_commit(&ctl.done_bar);
        }
    }
    // ---- epilogue: TMEM -> padded fp32 tile in shared memory (the pipeline stages are dead) -> coalesced 128-bit global accesses ------
    cp_async_wait<0>();
    if (ktiles > 0) {
        mbar_wait(&ctl.done_bar, 0);
        fence_after_sync();
    }
    gm_stamp(6);
    float* ct = reinterpret_cast<float*>(smem);
    {
        const int row = (warp & 3) * 32 + lane;                      // warps w and w + 4 share TMEM lane quadrant w & 3 and split the columns
        constexpr int CHUNKS8 = BN / 8, HALF = (CHUNKS8 + 1) / 2;
        const int c_beg = (warp >> 2) * HALF, c_end = min(CHUNKS8, c_beg + HALF);
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int CB = HALF < 4 ? HALF : 4;                      // chunks per batch: all loads of a batch in flight, one wait
        for (int c0 = c_beg; c0 < c_end; c0 += CB) {
            float v[CB][8];
#pragma unroll
            for (int j = 0; j < CB; ++j) {
                if (ktiles > 0 && c0 + j < c_end) tmem_ld_x8(tacc + (c0 + j) * 8, v[j]);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
                }
            }
            if (ktiles > 0) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CB; ++j)
                if (c0 + j < c_end) {
                    float4* d = reinterpret_cast<float4*>(ct + row * CF::CT_LD + (c0 + j) * 8);
                    d[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
                    d[1] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
                }
        }
    }
    fence_before_sync();
    __syncthreads();
    gm_stamp(7);
    const bool split = g.nsplit > 1;
    float* __restrict__ out = split ? g.partial + (size_t)blockIdx.z * g.M * g.N : g.C + (size_t)bz * g.sC;
    const int ldo = split ? g.N : g.ldc;
    const float* __restrict__ bias = (!split && g.bias) ? g.bias + (size_t)bz * g.sBias : nullptr;
    const float* __restrict__ Hb = (!split && g.H) ? g.H + (size_t)bz * g.sH : nullptr;
    const int act = split ? 0 : g.act, accumulate = split ? 0 : g.accumulate;
    // A thread owns ONE float4 column group (GM_THREADS is a multiple of the groups per row) and every RPI-th row of it: its bias values are
    // loaded once, and U rows are in flight together -- all shared / global loads of a batch are issued before the first dependent
    // instruction (the one-row-at-a-time loop spent 4.7 of the 8.5 us a 128-wide CTA of the LSTM heads lives in this phase: one L2 round
    // trip or one tanh chain per iteration, 2-4 warps per scheduler).
    constexpr int Q = BN / 4, RPI = GM_THREADS / Q, ITER = GM_BM / RPI, U = ITER < 4 ? ITER : 4;
    static_assert(GM_THREADS % Q == 0 && GM_BM % RPI == 0 && ITER % U == 0, "output mapping");
    const int q = tid % Q, r0 = tid / Q, n = n0 + q * 4;
    const int nv = min(4, g.N - n);                                  // <= 0: this thread's columns lie outside the matrix
    const bool full = nv == 4;
    // may rows of this output / bias / H be accessed as aligned float4 (n0 and the group offsets are multiples of 4)?
    const bool vec_o = full && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo & 3) == 0;
    const bool vec_b = full && bias && (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
    const bool vec_h = full && Hb && (reinterpret_cast<uintptr_t>(Hb) & 15) == 0 && (g.ldh & 3) == 0;
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias && nv > 0) {
        if (vec_b) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n)); b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w; }
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < nv) b[i] = __ldg(bias + n + i);
        }
    }
    if (nv > 0) {
        for (int it0 = 0; it0 < ITER; it0 += U) {
            float x[U][4], h[U][4], c[U][4];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                            // ---- loads of the batch
                const int r = r0 + (it0 + u) * RPI, m = m0 + r;
                ok[u] = m < g.M;
                const float4 cv = *reinterpret_cast<const float4*>(ct + r * CF::CT_LD + q * 4);
                x[u][0] = cv.x; x[u][1] = cv.y; x[u][2] = cv.z; x[u][3] = cv.w;
#pragma unroll
                for (int i = 0; i < 4; ++i) { h[u][i] = 0.f; c[u][i] = 0.f; }
                if (!ok[u]) continue;
                if (Hb) {
                    const float* hp = Hb + (size_t)m * g.ldh + n;
                    if (vec_h) { const float4 t = __ldg(reinterpret_cast<const float4*>(hp)); h[u][0] = t.x; h[u][1] = t.y; h[u][2] = t.z; h[u][3] = t.w; }
                    else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (i < nv) h[u][i] = __ldg(hp + i);
                    }
                }
                if (accumulate) {
                    const float* o = out + (size_t)m * ldo + n;
                    if (vec_o) { const float4 t = *reinterpret_cast<const float4*>(o); c[u][0] = t.x; c[u][1] = t.y; c[u][2] = t.z; c[u][3] = t.w; }
                    else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (i < nv) c[u][i] = o[i];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {                            // ---- epilogue arithmetic and stores
                if (!ok[u]) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = x[u][i];
                    if (bias) v += b[i];
                    if (act == 1) v = tanh_mufu(v);        // ex2 + rcp, |error| ~3e-7: tanhf's ~25 instructions made this phase issue-bound
                    if (Hb) v *= fmaf(-h[u][i], h[u][i], 1.f);
                    if (accumulate) v = c[u][i] + v;
                    x[u][i] = v;
                }
                float* o = out + (size_t)(m0 + r0 + (it0 + u) * RPI) * ldo + n;
                if (vec_o) *reinterpret_cast<float4*>(o) = make_float4(x[u][0], x[u][1], x[u][2], x[u][3]);
                else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (i < nv) o[i] = x[u][i];
                }
            }
        }
    }
    gm_stamp(8);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
    gm_stamp(9);
}

// C (+)= epilogue(sum over the split-K partial tiles, in slice order), for every batch entry
__global__ void k_gemm_splitk_reduce(const GemmArgs g, int nsplit, int batch) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, mn = (int64_t)g.M * g.N;
    if (idx >= mn * batch) return;
    const int bz = (int)(idx / mn);
    const int64_t r = idx - (int64_t)bz * mn;
    const int m = (int)(r / g.N), n = (int)(r - (int64_t)m * g.N);
    float x = 0.f;
    for (int z = 0; z < nsplit; ++z) x += g.partial[((size_t)bz * nsplit + z) * mn + r];
    if (g.bias) x += __ldg(g.bias + (size_t)bz * g.sBias + n);
    if (g.act == 1) x = tc::tanh_mufu(x);
    if (g.H) { const float h = __ldg(g.H + (size_t)bz * g.sH + (size_t)m * g.ldh + n); x *= fmaf(-h, h, 1.f); }
    float* c = g.C + (size_t)bz * g.sC + (size_t)m * g.ldc + n;
    if (g.accumulate) x += *c;
    *c = x;
}

template <int BN, bool PACKED> static int launch_gemm(const GemmArgs& g, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = GemmCfg<BN, PACKED>::SMEM;
    static bool seen[RB_MAX_DEVICES] = {};             // function attributes are per device
    if (first_use_on_device(seen)) RB_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3<BN, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_bf16x3<BN, PACKED><<<grid, GM_THREADS, smem, st>>>(g);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

template <int BN> static int launch_gemm(bool packed, const GemmArgs& g, dim3 grid, cudaStream_t st) {
    return packed ? launch_gemm<BN, true>(g, grid, st) : launch_gemm<BN, false>(g, grid, st);
}

// pipeline shape (GemmCfg): 0 = DEEP, 1 = PACKED, -1 = PACKED whenever the grid holds more CTAs than the GPU has SMs
static int g_gemm_cta_packing = -1;

int gemm_bf16x3(GemmArgs g, float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st) {
    RB_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0, "bad GEMM shape");
    if (g.batch <= 0) g.batch = 1;
    const int bn = g.N > 64 ? 128 : (g.N > 32 ? 64 : (g.N > 16 ? 32 : 16));
    const int tiles = ((g.M + GM_BM - 1) / GM_BM) * ((g.N + bn - 1) / bn) * g.batch;
    int split = 1;
    if (splitk_ws) {
        if (g.force_split > 0) split = g.force_split;
        // tiles * split <= 2 * sms: one wave of the BN <= 64 kernels (two CTAs per SM), two full waves of the <128> kernel -- rounding the
        // quotient UP left a nearly empty third wave (LSTM W_l gradient: 14 tiles x 22 slices = 308 CTAs on 148 SMs)
        else if (tiles < sms && g.K >= 8 * GM_BK) split = min(min(32, max(1, (2 * sms) / tiles)), g.K / (4 * GM_BK));
    }
    split = max(1, split);
    while (split > 1 && (size_t)split * g.batch * g.M * g.N > splitk_ws_floats) --split;
    int ks = ((g.K + split - 1) / split + GM_BK - 1) / GM_BK * GM_BK;
    if (ks == 0) ks = GM_BK;
    split = g.K > 0 ? (g.K + ks - 1) / ks : 1;
    g.ksplit = ks; g.nsplit = split; g.partial = splitk_ws;
    const dim3 grid((g.M + GM_BM - 1) / GM_BM, (g.N + bn - 1) / bn, split * g.batch);
    int rc;
    const bool packed = g_gemm_cta_packing < 0 ? (int64_t)grid.x * grid.y * grid.z > sms : g_gemm_cta_packing != 0;
    switch (bn) {
        case 128: rc = launch_gemm<128>(packed, g, grid, st); break;
        case 64: rc = launch_gemm<64>(packed, g, grid, st); break;
        case 32: rc = launch_gemm<32>(packed, g, grid, st); break;
        default: rc = launch_gemm<16>(packed, g, grid, st); break;
    }
    if (rc) return rc;
    if (split > 1) {
        const int64_t total = (int64_t)g.M * g.N * g.batch;
        k_gemm_splitk_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, split, g.batch);
        RB_CUDA(cudaGetLastError());
    }
    return RB_OK;
}

__global__ void k_sum_serial(const float* __restrict__ x, int n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < n; ++i) t += x[i];
        *out = t;
    }
}

// column sums (bias gradients) in two fixed-order stages: partial[bz][rb][col] over row blocks, then the sum over rb
__global__ void k_colsum_partial(const float* __restrict__ X, int ld, int64_t rows, int n, long long sX, int RB, float* __restrict__ partial) {
    __shared__ float red[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x, rb = blockIdx.y, bz = blockIdx.z;
    const int64_t chunk = (rows + RB - 1) / RB, r0 = rb * chunk, r1 = min(rows, r0 + chunk);
    const float* Xb = X + (size_t)bz * sX;
    float acc = 0.f;
    if (col < n)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += Xb[r * ld + col];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        partial[((size_t)bz * RB + rb) * n + col] = t;
    }
}
__global__ void k_colsum_final(const float* __restrict__ partial, int RB, int n, int batch, float* __restrict__ out, long long sOut) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * batch) return;
    const int bz = idx / n, col = idx - bz * n;
    float t = 0.f;
    for (int rb = 0; rb < RB; ++rb) t += partial[((size_t)bz * RB + rb) * n + col];
    out[(size_t)bz * sOut + col] = t;
}
int colsum(const float* X, int ld, int64_t rows, int n, int batch, long long sX, float* out, long long sOut, float* colpart, cudaStream_t st) {
    int RB = (int)min((int64_t)64, (rows + 63) / 64);
    while (RB > 1 && (size_t)RB * n * batch > COLPART_FLOATS) --RB;
    k_colsum_partial<<<dim3((n + 31) / 32, RB, batch), dim3(32, 8), 0, st>>>(X, ld, rows, n, sX, RB, colpart);
    k_colsum_final<<<(n * batch + 255) / 256, 256, 0, st>>>(colpart, RB, n, batch, out, sOut);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int sum_serial(const float* x, int n, float* out, cudaStream_t st) {
    k_sum_serial<<<1, 32, 0, st>>>(x, n, out);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // namespace rb

using namespace rb;

// debug: phase stamps (ns) of CTA 0 of the last k_gemm_bf16x3 launch: 0 start, 1 prologue done, 2 / 4 copies of k-tile 0 / 1 landed (and bf16 stage
// free), 3 / 5 converted + CTA barrier, 6 all MMAs done, 7 accumulator staged in shared memory, 8 output written, 9 end
extern "C" int rb_debug_gemm_stamps(unsigned long long* host_out) {
    RB_REQUIRE(host_out != nullptr, "NULL argument");
    RB_CUDA(cudaMemcpyFromSymbol(host_out, rb::g_gemm_stamps, sizeof(unsigned long long) * 16));
    return RB_OK;
}

extern "C" int rb_gemm_set_cta_packing(int mode) {
    RB_REQUIRE(mode >= -1 && mode <= 1, "mode must be -1 (by grid size), 0 (deep pipeline) or 1 (packed CTAs)");
    g_gemm_cta_packing = mode;
    return RB_OK;
}

// C[M,N] (+)= epilogue(A * B); see the header of this file for the operand orientation flags.  workspace (optional) enables split-K.
extern "C" int rb_gemm_bf16x3(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, float* C, int ldc, int M, int N, int K,
                              const float* bias, int act, int accumulate, const float* H, int ldh, float* workspace, int64_t workspace_floats,
                              void* stream) {
    RB_REQUIRE(A && B && C, "NULL argument");
    RB_REQUIRE((a_mn == 0 || a_mn == 1) && (b_mn == 0 || b_mn == 1) && (act == 0 || act == 1), "bad flag");
    int device = 0, sms = 148;
    RB_CUDA(cudaGetDevice(&device));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    GemmArgs g{};
    g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = B; g.ldb = ldb; g.b_mn = b_mn; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.act = act; g.accumulate = accumulate; g.H = H; g.ldh = ldh;
    return gemm_bf16x3(g, workspace, (size_t)(workspace_floats > 0 ? workspace_floats : 0), sms, (cudaStream_t)stream);
}
