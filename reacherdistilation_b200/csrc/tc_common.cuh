// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), shared by the tensor-core kernels.
//
// Operand layout used everywhere here: UMMA "no-swizzle, K-major" canonical layout.  An operand tile [R rows x K] of
// bf16 is stored as core matrices of 8 rows x 16 bytes (8 bf16 along K), each core matrix 128 contiguous bytes:
//     byte_offset(r, k) = (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2,   SBO = 128, LBO = R * 16
// i.e. [K/8][R][8]: a thread that owns row r writes one 16-byte chunk per 8 k-values at chunk*LBO + r*16 -- consecutive
// threads hit consecutive 16-byte slots (bank-conflict free 128-bit stores).  One tcgen05.mma kind::f16 consumes K = 16
// (two chunks); advancing K by 16 adds 2*LBO to the descriptor start address.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (64-bit): start address, leading (K-direction) and stride (row-group) byte offsets,
// descriptor version 1 (Blackwell), no swizzle.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// advance a descriptor's start address by `bytes` (multiple of 16; the 14-bit field cannot overflow for shared-memory addresses)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// One lane of a converged warp (elect.sync).  tcgen05.mma / commit issued under this predicate compile to a single UTCHMMA; under a
// per-thread condition such as `threadIdx.x == 0` ptxas wraps every MMA in an ELECT / BRA.U.ANY waterfall loop (~50 cycles each).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(mbar)), "r"(parity)
        : "memory");
}

// TMA bulk copy global -> shared (1-D), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM allocation: one full warp; the base address lands in *dst (shared memory)
template <int NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// TMEM -> registers, shape 32x32b: thread `lane` of the warp reads TMEM lane (warp%4)*32 + lane, N consecutive columns
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: ONE issue slot for two IEEE-rn fp32 operations).  A pair lives in a 64-bit
// register (slot 0 = low word).  ptxas folds scalar broadcasts ({x, x}), swapped halves ({hi, lo}) and negations into operand modifiers.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float s0, float s1) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(s0), "f"(s1)); return r; }
__device__ __forceinline__ void upk2(f32x2 r, float& s0, float& s1) { asm("mov.b64 {%0,%1}, %2;" : "=f"(s0), "=f"(s1) : "l"(r)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- bf16 hi/lo split (bf16x3 scheme): x ~= hi + lo, each bf16; two values packed per 32-bit word (first -> low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_half, float hi_half) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_half), "f"(lo_half));
    return d;
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi2, uint32_t& lo2) {
    hi2 = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hi2 << 16), bh = __uint_as_float(hi2 & 0xFFFF0000u);
    lo2 = pack_bf16x2(a - ah, b - bh);
}
__device__ __forceinline__ void split_scalar(float a, uint16_t& hi, uint16_t& lo) {
    uint32_t h2, l2;
    split_pair(a, 0.f, h2, l2);
    hi = (uint16_t)(h2 & 0xFFFFu);
    lo = (uint16_t)(l2 & 0xFFFFu);
}

// tanh(x) = 1 - 2 / (2^(2x log2 e) + 1) on the MUFU (ex2.approx + rcp.approx): absolute error ~3e-7
__device__ __forceinline__ float tanh_mufu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// same with the argument already multiplied by 2 log2(e) (the kernels fold that factor into the weights)
__device__ __forceinline__ float tanh_from_scaled(float y) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

// tanh of a PAIR from scaled arguments with ONE reciprocal: 1/(a b) shared, 1/a = b/(a b).  3 MUFU per pair instead of 4 (the XU
// pipe is the busiest pipe of the rollout kernel).  y is clamped to 60 so (e0+1)(e1+1) <= 2^121 stays finite; tanh(60/(2 log2 e)) == 1.
__device__ __forceinline__ void tanh_pair_from_scaled(float y0, float y1, float& t0, float& t1) {
    float e0, e1, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(y0, 60.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(y1, 60.f)));
    const float a0 = e0 + 1.0f, a1 = e1 + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a0 * a1));
    t0 = fmaf(-2.0f * a1, r, 1.0f);
    t1 = fmaf(-2.0f * a0, r, 1.0f);
}
// Packed form of tanh_pair_from_scaled: the +1 and the final 1 - 2 a r are one FADD2 / FFMA2 for the pair (same roundings, same
// bits).  Returns the pair as (tanh(y1'), tanh(y0')) i.e. SLOT 0 = t1, SLOT 1 = t0 (that is how 1 - 2 a_other r falls out of (a0, a1)).
__device__ __forceinline__ f32x2 tanh_pair_packed(float y0, float y1) {
    float e0, e1, r, a0, a1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(y0, 60.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(y1, 60.f)));
    const f32x2 a = add2(pk2(e0, e1), pk2(1.0f, 1.0f));
    upk2(a, a0, a1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a0 * a1));
    const float rr = -2.0f * r;
    return fma2(a, pk2(rr, rr), pk2(1.0f, 1.0f));
}
__device__ __forceinline__ void split_packed_swapped(f32x2 t, uint32_t& hi2, uint32_t& lo2);
// tanh pair + bf16 hi/lo split: 14 issue slots per pair (19 with scalar FADD / FFMA)
__device__ __forceinline__ void tanh_split_pair_packed(float y0, float y1, uint32_t& hi2, uint32_t& lo2) {
    split_packed_swapped(tanh_pair_packed(y0, y1), hi2, lo2);
}
// FOUR values with one reciprocal, packed: 5 MUFU and 18 issue slots per 4 elements (pair form: 6 MUFU, 18 slots).  y clamped to 30
// (tanh(30 / (2 log2 e)) rounds to 1.0f; the product of four (e + 1) <= 2^121 stays finite).  Returns (t1, t0) and (t3, t2).
__device__ __forceinline__ void tanh_quad_packed(const float* y, f32x2& t10, f32x2& t32) {
    float e[4], a0, a1, a2, a3, r;
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(fminf(y[i], 30.f)));
    const f32x2 a01 = add2(pk2(e[0], e[1]), pk2(1.0f, 1.0f)), a23 = add2(pk2(e[2], e[3]), pk2(1.0f, 1.0f));
    upk2(a01, a0, a1); upk2(a23, a2, a3);
    const float p01 = a0 * a1, p23 = a2 * a3;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float rr = -2.0f * r;
    float r01, r23;
    upk2(mul2(pk2(p23, p01), pk2(rr, rr)), r01, r23);      // -2 / (a0 a1), -2 / (a2 a3)
    t10 = fma2(a01, pk2(r01, r01), pk2(1.0f, 1.0f));
    t32 = fma2(a23, pk2(r23, r23), pk2(1.0f, 1.0f));
}
// bf16 hi/lo split of a packed pair given as (x1, x0): hi2 / lo2 hold element 0 in the low half
__device__ __forceinline__ void split_packed_swapped(f32x2 t, uint32_t& hi2, uint32_t& lo2) {
    float t1, t0, l1, l0;
    upk2(t, t1, t0);
    hi2 = pack_bf16x2(t0, t1);
    const float ah = __uint_as_float(hi2 << 16), bh = __uint_as_float(hi2 & 0xFFFF0000u);
    upk2(add2(t, pk2(-bh, -ah)), l1, l0);
    lo2 = pack_bf16x2(l0, l1);
}
// same for FOUR values with one reciprocal (5 MUFU per 4 elements); y clamped to 30: tanh(30 / (2 log2 e)) rounds to 1.0f exactly and
// the product of four (e + 1) <= 2^121 stays finite.
__device__ __forceinline__ void tanh_quad_from_scaled(const float* y, float* t) {
    float e[4], a[4], r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(fminf(y[i], 30.f)));
        a[i] = e[i] + 1.0f;
    }
    const float p01 = a[0] * a[1], p23 = a[2] * a[3];
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float r01 = -2.0f * (r * p23), r23 = -2.0f * (r * p01);      // -2 / (a0 a1), -2 / (a2 a3)
    t[0] = fmaf(r01, a[1], 1.0f);
    t[1] = fmaf(r01, a[0], 1.0f);
    t[2] = fmaf(r23, a[3], 1.0f);
    t[3] = fmaf(r23, a[2], 1.0f);
}
// variant with the reciprocal on the FMA pipe (integer seed + 3 Newton steps, rel. error < 1e-7): 1 MUFU per element
__device__ __forceinline__ float tanh_from_scaled_newton(float y) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(y, 100.f)));
    const float a = e + 1.0f;
    float r = __uint_as_float(0x7EF311C7u - __float_as_uint(a));
    r = r * fmaf(-a, r, 2.0f);
    r = r * fmaf(-a, r, 2.0f);
    r = r * fmaf(-a, r, 2.0f);
    return fmaf(-2.0f, r, 1.0f);
}

// byte offset of element (r, k) inside a K-major no-swizzle tile with R rows
__host__ __device__ constexpr uint32_t tile_off(int r, int k, int R) { return (uint32_t)((k >> 3) * (R * 16) + r * 16 + (k & 7) * 2); }

}  // namespace tc
}  // namespace rb
