// Counter-based synthetic R1CS recipe, device side (measurement fixture; DESIGN.md "Synthetic instances").
// CPU twins: oracle/synth.py (specification) and oracle/bp_oracle.c (bulk).  Written independently of both;
// tests/test_gpu_synth.py checks the three agree element for element.
#pragma once
#include <stdint.h>
#include "field.cuh"

namespace bp {

BP_HD uint64_t sm_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

BP_HD uint64_t sm_key(uint64_t seed, uint64_t stream, uint64_t i, uint64_t j) {
    return sm_mix(sm_mix(sm_mix(seed ^ (stream << 56)) + i) + j);
}

// Uniform element of [0, p) by rejection from 255 random bits; out = 8 x u32 limbs.
template <int F> BP_HD void sm_sample(uint64_t h, uint32_t* out) {
    uint64_t l[4];
    for (int a = 0; a < 64; ++a) {
#pragma unroll
        for (int k = 0; k < 4; ++k) l[k] = sm_mix(h + 4 * (uint64_t)a + (uint64_t)k + 1);
        l[3] &= 0x7fffffffffffffffULL;
#pragma unroll
        for (int k = 0; k < 4; ++k) { out[2 * k] = (uint32_t)l[k]; out[2 * k + 1] = (uint32_t)(l[k] >> 32); }
        if (is_canonical<F>(out)) return;
    }
    out[7] >>= 2;
}

BP_HD uint32_t sm_len(uint64_t seed, uint32_t t, uint64_t lcid) {
    return 1u + (uint32_t)(sm_key(seed, 1, lcid, 0) % (2ull * t - 1ull));
}

// k-th column of an LC of length `len`: stratified -> strictly ascending, unique.  Returns the tagged column.
BP_HD uint32_t sm_col(uint64_t seed, uint64_t lcid, uint32_t k, uint32_t len, uint64_t n_vars, uint64_t n_inputs) {
    const uint64_t lo = ((uint64_t)k * n_vars) / len, hi = ((uint64_t)(k + 1) * n_vars) / len;
    const uint64_t col = lo + sm_key(seed, 2, lcid, k) % (hi - lo);
    return col < n_inputs ? (uint32_t)col : ((uint32_t)(col - n_inputs) | 0x80000000u);
}

}  // namespace bp
