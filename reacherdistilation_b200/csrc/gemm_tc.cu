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
// 128 x BN output tile per CTA, BK = 32 per stage, 3 stages; accumulator in TMEM; MMAs issued by one thread, stage reuse gated by
// tcgen05.commit -> mbarrier.  Split-K (grid.z) writes partial tiles that a second kernel adds in a fixed order (deterministic).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace rb {

using namespace tc;

constexpr int GM_THREADS = 256;
constexpr int GM_BM = 128;
constexpr int GM_BK = 32;
constexpr int GM_STAGES = 3;

struct __align__(16) GemmCtl {
    uint64_t stage_bar[GM_STAGES];
    uint64_t done_bar;
    uint32_t tmem_base;
};

// one 8-float chunk of an operand tile: rows [row0, row0+R), k in [k0, k0+32); returns zeros outside [nrows) x [kend)
__device__ __forceinline__ void load_chunk(const float* __restrict__ X, int ld, int mn, int row, int k, int nrows, int kend, float* v) {
    const bool vec_ok = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    if (mn == 0) {                                   // 8 consecutive k of one row
        const float* p = X + (size_t)row * ld + k;
        if (row < nrows && k + 8 <= kend && vec_ok && (k & 3) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (row < nrows && k + i < kend) ? __ldg(p + i) : 0.f;
        }
    } else {                                         // 8 consecutive rows of one k
        const float* p = X + (size_t)k * ld + row;
        if (k < kend && row + 8 <= nrows && vec_ok && (row & 3) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (k < kend && row + i < nrows) ? __ldg(p + i) : 0.f;
        }
    }
}
__device__ __forceinline__ void store_chunk(uint8_t* hi, uint8_t* lo, uint32_t off, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pair(v[2 * q], v[2 * q + 1], h[q], l[q]);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Tile geometry of an operand with R rows (R = 128 for A, BN for B), BK = 32:
//   K-major : chunk (row r, k-group kg)  at kg * (R*16) + r*16          LBO = R*16, SBO = 128, K-step advance = 2 * LBO
//   MN-major: chunk (k, row-group rg)    at rg * 512 + (k/8)*128 + (k%8)*16     LBO = 128, SBO = 512, K-step advance = 256
template <int R> struct OperandTile {
    static constexpr int CHUNKS = R * GM_BK / 8;
    static constexpr int PER_THREAD = (CHUNKS + GM_THREADS - 1) / GM_THREADS;
    static constexpr int BYTES = R * GM_BK * 2;
    // chunk index c -> (row, k) of its first element and byte offset in the tile
    __device__ static __forceinline__ void map(int c, int mn, int& row, int& k, uint32_t& off) {
        if (mn == 0) { row = c % R; const int kg = c / R; k = kg * 8; off = kg * (R * 16) + row * 16; }
        else { k = c % GM_BK; const int rg = c / GM_BK; row = rg * 8; off = rg * 512 + (k >> 3) * 128 + (k & 7) * 16; }
    }
    __device__ static __forceinline__ uint64_t desc(uint32_t base, int mn, int ks) {
        return mn == 0 ? make_smem_desc(base + ks * 2 * (R * 16), R * 16, 128) : make_smem_desc(base + ks * 256, 128, 512);
    }
};

template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 2) k_gemm_bf16x3(const GemmArgs g) {
    using TA = OperandTile<GM_BM>;
    using TB = OperandTile<BN>;
    constexpr int STAGE_BYTES = 2 * TA::BYTES + 2 * TB::BYTES;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(128) uint8_t smem[];
    GemmCtl& ctl = *reinterpret_cast<GemmCtl*>(smem + GM_STAGES * STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * BN;
    const int bz = blockIdx.z / g.nsplit, sz = blockIdx.z - bz * g.nsplit;          // batch index, K slice
    const float* __restrict__ Ab = g.A + (size_t)bz * g.sA;
    const float* __restrict__ Bb = g.B + (size_t)bz * g.sB;
    const int kbeg = sz * g.ksplit, kend = min(g.K, kbeg + g.ksplit);
    const int ktiles = (kend - kbeg + GM_BK - 1) / GM_BK;

    if (warp == 0) tmem_alloc<TCOLS>(&ctl.tmem_base);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GM_STAGES; ++s) mbar_init(&ctl.stage_bar[s], 1);
        mbar_init(&ctl.done_bar, 1);
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = ctl.tmem_base;
    const uint32_t idesc = make_idesc_bf16(GM_BM, BN) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16);

    float va[TA::PER_THREAD][8], vb[TB::PER_THREAD][8];
    auto gload = [&](int kt) {
        const int k0 = kbeg + kt * GM_BK;
#pragma unroll
        for (int i = 0; i < TA::PER_THREAD; ++i) {
            const int c = tid + i * GM_THREADS;
            int row, k; uint32_t off;
            TA::map(c, g.a_mn, row, k, off);
            if (c < TA::CHUNKS) load_chunk(Ab, g.lda, g.a_mn, m0 + row, k0 + k, g.M, kend, va[i]);
        }
#pragma unroll
        for (int i = 0; i < TB::PER_THREAD; ++i) {
            const int c = tid + i * GM_THREADS;
            int row, k; uint32_t off;
            TB::map(c, g.b_mn, row, k, off);
            if (c < TB::CHUNKS) load_chunk(Bb, g.ldb, g.b_mn, n0 + row, k0 + k, g.N, kend, vb[i]);
        }
    };
    uint32_t stage_phase = 0;           // bit s = parity to wait for on stage_bar[s]
    if (ktiles > 0) gload(0);
    for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % GM_STAGES;
        uint8_t* a_hi = smem + s * STAGE_BYTES;
        uint8_t* a_lo = a_hi + TA::BYTES;
        uint8_t* b_hi = a_lo + TA::BYTES;
        uint8_t* b_lo = b_hi + TB::BYTES;
        if (kt >= GM_STAGES) {                                        // the MMAs that read this stage (tile kt - STAGES) are done
            mbar_wait(&ctl.stage_bar[s], (stage_phase >> s) & 1u);
            stage_phase ^= 1u << s;
        }
#pragma unroll
        for (int i = 0; i < TA::PER_THREAD; ++i) {
            const int c = tid + i * GM_THREADS;
            int row, k; uint32_t off;
            TA::map(c, g.a_mn, row, k, off);
            if (c < TA::CHUNKS) store_chunk(a_hi, a_lo, off, va[i]);
        }
#pragma unroll
        for (int i = 0; i < TB::PER_THREAD; ++i) {
            const int c = tid + i * GM_THREADS;
            int row, k; uint32_t off;
            TB::map(c, g.b_mn, row, k, off);
            if (c < TB::CHUNKS) store_chunk(b_hi, b_lo, off, vb[i]);
        }
        if (kt + 1 < ktiles) gload(kt + 1);                           // next tile's global loads fly while this tile's MMAs are issued
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
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
    // ---- epilogue: warps w and w+4 share TMEM lane quadrant w & 3 and split the columns ---------------------------------------
    if (ktiles > 0) {
        mbar_wait(&ctl.done_bar, 0);
        fence_after_sync();
    }
    const int row = (warp & 3) * 32 + lane, m = m0 + row;
    constexpr int CHUNKS8 = BN / 8, HALF = (CHUNKS8 + 1) / 2;
    const int c_beg = (warp >> 2) * HALF, c_end = min(CHUNKS8, c_beg + HALF);
    const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const bool split = g.nsplit > 1;
    float* out = split ? g.partial + (size_t)blockIdx.z * g.M * g.N : g.C + (size_t)bz * g.sC;
    const int ldo = split ? g.N : g.ldc;
    const float* bias = g.bias ? g.bias + (size_t)bz * g.sBias : nullptr;
    const float* Hb = g.H ? g.H + (size_t)bz * g.sH : nullptr;
    for (int c = c_beg; c < c_end; ++c) {
        float v[8];
        if (ktiles > 0) { tmem_ld_x8(tacc + c * 8, v); tmem_ld_wait(); }
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        if (m < g.M) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = n0 + c * 8 + i;
                if (n < g.N) {
                    float x = v[i];
                    if (!split) {
                        if (bias) x += __ldg(bias + n);
                        if (g.act == 1) x = tanhf(x);
                        if (Hb) { const float h = __ldg(Hb + (size_t)m * g.ldh + n); x *= fmaf(-h, h, 1.f); }
                        if (g.accumulate) x += out[(size_t)m * ldo + n];
                    }
                    out[(size_t)m * ldo + n] = x;
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
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
    if (g.act == 1) x = tanhf(x);
    if (g.H) { const float h = __ldg(g.H + (size_t)bz * g.sH + (size_t)m * g.ldh + n); x *= fmaf(-h, h, 1.f); }
    float* c = g.C + (size_t)bz * g.sC + (size_t)m * g.ldc + n;
    if (g.accumulate) x += *c;
    *c = x;
}

template <int BN> static int launch_gemm(const GemmArgs& g, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = GM_STAGES * (2 * OperandTile<GM_BM>::BYTES + 2 * OperandTile<BN>::BYTES) + sizeof(GemmCtl);
    static bool attr = false;
    if (!attr) { RB_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_gemm_bf16x3<BN><<<grid, GM_THREADS, smem, st>>>(g);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int gemm_bf16x3(GemmArgs g, float* splitk_ws, size_t splitk_ws_floats, int sms, cudaStream_t st) {
    RB_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0, "bad GEMM shape");
    if (g.batch <= 0) g.batch = 1;
    const int bn = g.N > 64 ? 128 : (g.N > 32 ? 64 : (g.N > 16 ? 32 : 16));
    const int tiles = ((g.M + GM_BM - 1) / GM_BM) * ((g.N + bn - 1) / bn) * g.batch;
    int split = 1;
    if (splitk_ws) {
        if (g.force_split > 0) split = g.force_split;
        else if (tiles < sms && g.K >= 8 * GM_BK) split = min(min(32, (2 * sms + tiles - 1) / tiles), g.K / (4 * GM_BK));
    }
    split = max(1, split);
    while (split > 1 && (size_t)split * g.batch * g.M * g.N > splitk_ws_floats) --split;
    int ks = ((g.K + split - 1) / split + GM_BK - 1) / GM_BK * GM_BK;
    if (ks == 0) ks = GM_BK;
    split = g.K > 0 ? (g.K + ks - 1) / ks : 1;
    g.ksplit = ks; g.nsplit = split; g.partial = splitk_ws;
    const dim3 grid((g.M + GM_BM - 1) / GM_BM, (g.N + bn - 1) / bn, split * g.batch);
    int rc;
    switch (bn) {
        case 128: rc = launch_gemm<128>(g, grid, st); break;
        case 64: rc = launch_gemm<64>(g, grid, st); break;
        case 32: rc = launch_gemm<32>(g, grid, st); break;
        default: rc = launch_gemm<16>(g, grid, st); break;
    }
    if (rc) return rc;
    if (split > 1) {
        const int64_t total = (int64_t)g.M * g.N * g.batch;
        k_gemm_splitk_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, split, g.batch);
        RB_CUDA(cudaGetLastError());
    }
    return RB_OK;
}

}  // namespace rb

using namespace rb;

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
