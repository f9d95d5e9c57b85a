// Threefry-2x32-20 and the jax.random derivations the NUTS transition consumes
// (split / random_bits / uniform / bernoulli / normal, partitionable mode).
// Replaces the jax.random call sites listed in SURVEY.md row a14 (numpyro/infer/hmc.py:94,102,
// 335,472-474,745-750; hmc_util.py:355,565,656,804,920,1005,1161-1162; infer/util.py:423,460-463;
// mcmc.py:671).  Device twin of oracle/prng.py; bit-exact by construction (integer arithmetic +
// det-f32 float transforms).
#pragma once
#include "common.cuh"
#include "detmath.cuh"

namespace b2 {

struct Key { uint32_t a, b; };

B2_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

B2_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    uint32_t x0 = c0 + k0, x1 = c1 + k1;
#define B2_TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
    B2_TF_R(13) B2_TF_R(15) B2_TF_R(26) B2_TF_R(6)
    x0 += k1; x1 += k2 + 1u;
    B2_TF_R(17) B2_TF_R(29) B2_TF_R(16) B2_TF_R(24)
    x0 += k2; x1 += k0 + 2u;
    B2_TF_R(13) B2_TF_R(15) B2_TF_R(26) B2_TF_R(6)
    x0 += k0; x1 += k1 + 3u;
    B2_TF_R(17) B2_TF_R(29) B2_TF_R(16) B2_TF_R(24)
    x0 += k1; x1 += k2 + 4u;
    B2_TF_R(13) B2_TF_R(15) B2_TF_R(26) B2_TF_R(6)
    x0 += k2; x1 += k0 + 5u;
#undef B2_TF_R
    o0 = x0; o1 = x1;
}

// jax.random.split(key, n)[i]  (partitionable: counter (0, i), both output words)
B2_HD Key split_at(Key k, uint32_t i) {
    Key o; threefry2x32(k.a, k.b, 0u, i, o.a, o.b); return o;
}

// jax.random.bits(key, shape)[i] for a flat index i < 2^32
B2_HD uint32_t bits_at(Key k, uint32_t i) {
    uint32_t x0, x1; threefry2x32(k.a, k.b, 0u, i, x0, x1); return x0 ^ x1;
}

B2_HD float unit_from_bits(uint32_t bits) { return bits_to_float((bits >> 9) | 0x3F800000u) - 1.0f; }

// jax.random.uniform(key, shape)[i] in [0, 1)
B2_HD float uniform01_at(Key k, uint32_t i) { return unit_from_bits(bits_at(k, i)); }

// jax.random.uniform(key, shape, minval=lo, maxval=hi)[i] = max(lo, u * (hi - lo) + lo)
B2_HD float uniform_at(Key k, uint32_t i, float lo, float hi) {
    const float u = unit_from_bits(bits_at(k, i));
    const float v = u * (hi - lo) + lo;
    return (lo > v) ? lo : v;
}

// jax.random.normal(key, shape)[i] = sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))
B2_HD float normal_at(Key k, uint32_t i) {
    const float lo = bits_to_float(0xBF7FFFFFu);     // nextafter(-1, 0)
    const float u = uniform_at(k, i, lo, 1.0f);
    return 1.41421356237309505f * d_erfinv(u);
}

}  // namespace b2
