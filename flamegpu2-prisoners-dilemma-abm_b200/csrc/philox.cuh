// Philox4x32-10 (Salmon et al., SC'11) as one inlined device function.
//
// Replaces FLAMEGPU2's per-thread cuRAND streams behind `FLAMEGPU->random.*`
// (/root/reference/src/model.py:364, :611-612, :631, :664, :806-810, :1076-1079, :1236,
// :1263-1314): every draw is the word of a counter (cell_lo, cell_hi, step, stream) under
// the key (seed_lo, seed_hi), so a value depends on nothing but WHO draws WHEN -- not on
// launch geometry, list order or the number of GPUs.  Same constants as cuRAND's
// curand_philox4x32_x.h; checked against the Random123 known-answer vectors in
// tests/test_gpu_parity.py::test_philox_kat.
#pragma once
#include <stdint.h>

namespace pd {

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                      uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}

struct Words { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Words philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Words w;
    w.x = c0; w.y = c1; w.z = c2; w.w = c3;
    return w;
}

}  // namespace pd
