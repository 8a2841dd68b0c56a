// Philox4x32-10 counter-based RNG (Salmon et al. 2011) -- identical arithmetic to oracle/philox_np.py and
// oracle/reacher_oracle.c, so resets / actions / dropout masks are bit-exact between host oracle and device.
// counter = (global env or sample id, episode or step or iteration, draw index, stream id), key = (seed_lo, seed_hi).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rb {

enum : uint32_t { STREAM_RESET = 0u, STREAM_ACTION = 1u, STREAM_DROPOUT = 2u };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                        uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// u = (x >> 8) * 2^-24 in [0,1); one fused multiply-add maps it to [lo, hi).
__host__ __device__ __forceinline__ float uniform_f32(uint32_t x, float lo, float hi) {
    float u = (float)(x >> 8) * 5.9604644775390625e-08f;
#ifdef __CUDA_ARCH__
    return __fmaf_rn(u, hi - lo, lo);
#else
    return fmaf(u, hi - lo, lo);
#endif
}

}  // namespace rb
