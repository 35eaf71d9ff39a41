// mtg_philox.cuh -- Philox4x32-10 counter-based generator (Salmon et al. 2011): the random source of every
// device-side draw (simulator spectrum / segment / noise, ensemble sampler).  Integer arithmetic only, so the
// stream does not depend on the floating-point contraction mode of the including translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double u01_from_bits(uint32_t a, uint32_t b)      // (0, 1), 53 bits
{
    const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
    return ((double)v + 0.5) * 1.1102230246251565e-16;
}
