// qz_philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), host + device.
// Stream layout used by the engine (DESIGN.md "rollout sampling"):
//   key = 64-bit seed;  counter = (rid_lo, rid_hi, c2, c3) with rid a globally unique rollout / game id,
//   so results do not depend on how games are sharded over GPUs.
#pragma once
#include <stdint.h>

#include "qz_rules.cuh"

struct QzPhilox4 {
    uint32_t x, y, z, w;
};

QZ_HD uint32_t qz_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

QZ_HD QzPhilox4 qz_philox(uint64_t seed, uint64_t rid, uint32_t c2, uint32_t c3) {
    uint32_t c0 = (uint32_t)rid, c1 = (uint32_t)(rid >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = qz_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = qz_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    QzPhilox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

QZ_HD uint32_t qz_philox_word(const QzPhilox4 &b, int i) { return i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w)); }
