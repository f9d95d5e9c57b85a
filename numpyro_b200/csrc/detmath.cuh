// det-f32 transcendental functions: device twin of oracle/detmath.py (same constants, same
// operation order, one rounding per operation; this header must be compiled without FMA
// contraction).  Used by the tree / adaptation bookkeeping so that acceptance decisions, step
// sizes and Gaussian momenta are bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace b2 {

B2_HD float pow2i(int k) { return bits_to_float((uint32_t)(k + 127) << 23); }   // -126 <= k <= 127

B2_HD float d_exp(float x) {
    if (is_nan(x)) return x;
    if (x > 88.72283935546875f) return f_inf();
    if (x < -103.972084045410f) return 0.0f;
    const float t = x * 1.44269504088896341f;
    const float half = (t >= 0.0f) ? 0.5f : -0.5f;
    const int k = (int)(t + half);                 // truncation toward zero
    const float kf = (float)k;
    const float r = (x - kf * 0.693359375f) - kf * -2.12194440e-4f;
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    const float z = r * r;
    const float y = (p * z + r) + 1.0f;
    const int k1 = k / 2;
    const int k2 = k - k1;
    return (y * pow2i(k1)) * pow2i(k2);
}

B2_HD float d_log(float x) {
    if (is_nan(x) || x < 0.0f) return f_nan();
    if (x == 0.0f) return -f_inf();
    if (is_inf(x)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    const uint32_t b = float_to_bits(x);
    e += (int)((b >> 23) & 0xFFu) - 126;
    float m = bits_to_float((b & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    const float ef = (float)e;
    const float z = m * m;
    float p = 7.0376836292e-2f;
    p = p * m + -1.1514610310e-1f;
    p = p * m + 1.1676998740e-1f;
    p = p * m + -1.2420140846e-1f;
    p = p * m + 1.4249322787e-1f;
    p = p * m + -1.6668057665e-1f;
    p = p * m + 2.0000714765e-1f;
    p = p * m + -2.4999993993e-1f;
    p = p * m + 3.3333331174e-1f;
    float y = (p * m) * z;
    y = y + -2.12194440e-4f * ef;
    y = y - 0.5f * z;
    const float r = m + y;
    return r + 0.693359375f * ef;
}

B2_HD float d_log1p(float x) {
    if (is_nan(x)) return x;
    const float u = 1.0f + x;
    if (u == 1.0f) return x;
    if (is_inf(u)) return u;
    return (d_log(u) * x) / (u - 1.0f);
}

B2_HD float d_expit(float x) { return 1.0f / (1.0f + d_exp(-x)); }

B2_HD float d_logaddexp(float a, float b) {
    const float d = a - b;
    if (is_nan(d)) return a + b;
    const float amax = (a >= b) ? a : b;
    return amax + d_log1p(d_exp(-fabsf(d)));
}

B2_HD float d_pow(float x, float y) { return d_exp(y * d_log(x)); }

B2_HD float d_erfinv(float x) {
    if (fabsf(x) == 1.0f) return (x > 0.0f) ? f_inf() : -f_inf();
    float w = -d_log1p(-(x * x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = 3.43273939e-07f + p * w;
        p = -3.5233877e-06f + p * w;
        p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w;
        p = -0.00125372503f + p * w;
        p = -0.00417768164f + p * w;
        p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = 0.000100950558f + p * w;
        p = 0.00134934322f + p * w;
        p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w;
        p = -0.0076224613f + p * w;
        p = 0.00943887047f + p * w;
        p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}

}  // namespace b2
