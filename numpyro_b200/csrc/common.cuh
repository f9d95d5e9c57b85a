// Shared host/device plumbing for the NUTS engine.
//
// The per-chain state machine (tick.cuh) is single-source: nvcc compiles it for sm_100a where a
// chain is owned by one warp and a D-vector is strided across the 32 lanes; g++ compiles the very
// same text for the test-only host simulator (tests/hostsim), where "a warp" is a plain loop.
// Both follow the det-f32 convention of oracle/detmath.py: one IEEE binary32 rounding per written
// operation, no FMA contraction (this TU is built with -fmad=false / -ffp-contract=off), and the
// lane-strided + xor-butterfly reduction order.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#define B2_HD_COLD static __host__ __device__ __noinline__      // rarely taken paths: kept out of line (code size, build time)
#else
#define B2_HD inline
#define B2_D inline
#define B2_HD_COLD inline
#endif

namespace b2 {

B2_HD float bits_to_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
B2_HD uint32_t float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4); return b;
#endif
}
B2_HD bool is_nan(float x) { return x != x; }
B2_HD bool is_inf(float x) { return (float_to_bits(x) & 0x7FFFFFFFu) == 0x7F800000u; }
B2_HD bool is_finite(float x) { return (float_to_bits(x) & 0x7F800000u) != 0x7F800000u; }
B2_HD float f_inf() { return bits_to_float(0x7F800000u); }
B2_HD float f_nan() { return bits_to_float(0x7FC00000u); }
B2_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// ---- lane model -----------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
B2_D int lane_first() { return (int)(threadIdx.x & 31u); }
B2_D int lane_step() { return 32; }
#else
inline int lane_first() { return 0; }
inline int lane_step() { return 1; }
#endif

#define B2_FOR_D(d, D) for (int d = b2::lane_first(); d < (D); d += b2::lane_step())

// Warp barrier between lanes that hand vector elements to each other through memory.
#if defined(__CUDA_ARCH__)
// A shuffle whose result is consumed, not __syncwarp(): ptxas 12.9 was seen to turn a __syncwarp() that follows a
// lane-strided loop into a NOP (stream_engine.cuh, gred reduction); a shuffle cannot be dropped.
B2_D void lane_sync() { unsigned int x = __shfl_sync(0xFFFFFFFFu, threadIdx.x, 0); asm volatile("" ::"r"(x) : "memory"); }
#else
inline void lane_sync() {}
#endif

// Canonical reduction of f(d), d in [0, D): lane partials p[l] = sum_k f(l + 32k) (k ascending,
// starting from +0), then p[l] += p[l ^ off] for off = 16, 8, 4, 2, 1.  All lanes get the result.
template <class Fn>
B2_HD float lane_sum(int D, Fn f) {
#if defined(__CUDA_ARCH__)
    float p = 0.0f;
    for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) p = p + f(d);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = p + __shfl_xor_sync(0xFFFFFFFFu, p, off);
    return p;
#else
    float p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0f;
    for (int d = 0; d < D; ++d) p[d & 31] = p[d & 31] + f(d);
    for (int off = 16; off > 0; off >>= 1) {
        float q[32];
        for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; ++l) p[l] = q[l];
    }
    return p[0];
#endif
}

// Two reductions sharing one butterfly (same per-value order as two lane_sum calls).
template <class Fn>
B2_HD void lane_sum2(int D, Fn f, float& out0, float& out1) {
#if defined(__CUDA_ARCH__)
    float p0 = 0.0f, p1 = 0.0f;
    for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) {
        float a, b; f(d, a, b); p0 = p0 + a; p1 = p1 + b;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        p0 = p0 + __shfl_xor_sync(0xFFFFFFFFu, p0, off);
        p1 = p1 + __shfl_xor_sync(0xFFFFFFFFu, p1, off);
    }
    out0 = p0; out1 = p1;
#else
    float p[32], r[32];
    for (int l = 0; l < 32; ++l) { p[l] = 0.0f; r[l] = 0.0f; }
    for (int d = 0; d < D; ++d) { float a, b; f(d, a, b); p[d & 31] = p[d & 31] + a; r[d & 31] = r[d & 31] + b; }
    for (int off = 16; off > 0; off >>= 1) {
        float q[32], s[32];
        for (int l = 0; l < 32; ++l) { q[l] = p[l] + p[l ^ off]; s[l] = r[l] + r[l ^ off]; }
        for (int l = 0; l < 32; ++l) { p[l] = q[l]; r[l] = s[l]; }
    }
    out0 = p[0]; out1 = r[0];
#endif
}

// ---- wide vectors (D > 64: the GEMM regime's few-but-wide chains) ---------------------------------------------------------------
// A lane's element loop is a chain of dependent load -> compute -> store round trips (~600 cycles each from L2 / HBM) and a warp
// owns one chain, so nothing hides the latency.  These variants keep FOUR elements of a lane in flight: all loads of a batch are
// issued before its first store (explicitly, so no alias analysis is needed).  Per element the arithmetic is unchanged and the
// reductions add in the same order (k ascending per lane, then the butterfly): identical bits.
struct Vals { float x[6]; };
template <class L, class S>
B2_HD void for_d_wide(int D, L load, S store) {
#if defined(__CUDA_ARCH__) && defined(B2_NO_WIDE)
    // (translation units whose chain vectors live in shared memory -- the streaming regime -- keep the plain loops: nothing to
    //  hide there, and the extra code slowed that kernel's latency-critical tick through its instruction-cache footprint)
    for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) { const Vals a = load(d); store(d, a); }
#elif defined(__CUDA_ARCH__)
    int d = (int)(threadIdx.x & 31u);
    for (; d + 96 < D; d += 128) {
        const Vals a = load(d), b = load(d + 32), c = load(d + 64), e = load(d + 96);
        store(d, a); store(d + 32, b); store(d + 64, c); store(d + 96, e);
    }
    for (; d < D; d += 32) { const Vals a = load(d); store(d, a); }
#else
    for (int d = 0; d < D; ++d) { const Vals a = load(d); store(d, a); }
#endif
}
template <class Fn>
B2_HD float lane_sum_wide(int D, Fn f) {
#if defined(__CUDA_ARCH__) && !defined(B2_NO_WIDE)
    float p = 0.0f;
    int d = (int)(threadIdx.x & 31u);
    for (; d + 96 < D; d += 128) {
        const float a = f(d), b = f(d + 32), c = f(d + 64), e = f(d + 96);
        p = p + a; p = p + b; p = p + c; p = p + e;
    }
    for (; d < D; d += 32) p = p + f(d);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = p + __shfl_xor_sync(0xFFFFFFFFu, p, off);
    return p;
#else
    return lane_sum(D, f);
#endif
}
template <class Fn>
B2_HD void lane_sum2_wide(int D, Fn f, float& out0, float& out1) {
#if defined(__CUDA_ARCH__) && !defined(B2_NO_WIDE)
    float p0 = 0.0f, p1 = 0.0f;
    int d = (int)(threadIdx.x & 31u);
    for (; d + 96 < D; d += 128) {
        float a0, b0, a1, b1, a2, b2, a3, b3;
        f(d, a0, b0); f(d + 32, a1, b1); f(d + 64, a2, b2); f(d + 96, a3, b3);
        p0 = p0 + a0; p1 = p1 + b0; p0 = p0 + a1; p1 = p1 + b1; p0 = p0 + a2; p1 = p1 + b2; p0 = p0 + a3; p1 = p1 + b3;
    }
    for (; d < D; d += 32) { float a, b; f(d, a, b); p0 = p0 + a; p1 = p1 + b; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        p0 = p0 + __shfl_xor_sync(0xFFFFFFFFu, p0, off);
        p1 = p1 + __shfl_xor_sync(0xFFFFFFFFu, p1, off);
    }
    out0 = p0; out1 = p1;
#else
    lane_sum2(D, f, out0, out1);
#endif
}

// lane-wide "any": a per-lane flag OR-reduced over the lanes owning d in [0, D)
template <class Fn>
B2_HD bool lane_any(int D, Fn f) {
#if defined(__CUDA_ARCH__)
    bool v = false;
    for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) v = v || f(d);
    return __any_sync(0xFFFFFFFFu, v);
#else
    bool v = false;
    for (int d = 0; d < D; ++d) v = v || f(d);
    return v;
#endif
}

}  // namespace b2
