// Per-observation link functions of the GLM likelihood sweeps (streaming regime R2 and GEMM regime R3).
#pragma once
#include "tick.cuh"        // (defines B2_LAPQ, which families.cuh uses)
#include "families.cuh"

namespace b2 {

B2_D float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
B2_D float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
B2_D float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Per-observation loss and d loss / d eta with the hardware ex2/lg2/rcp approximations (relative
// error ~1e-7 per term, far inside the rtol 1e-5 parity budget once summed over rows).  Same
// formulas as glm_loss (families.cuh): distributions/util.py:317-320, discrete.py:1388.
template <int LIK>
B2_D void link_fn(float eta, float y, float& loss, float& dl) {
    if (LIK == LIK_BERNOULLI) {
        const float e = ex2_approx(fabsf(eta) * -1.4426950408889634f);       // exp(-|eta|), in (0, 1]
        const float ope = 1.0f + e;
        loss = __fmaf_rn(lg2_approx(ope), 0.6931471805599453f, __fmaf_rn(-eta, y, fmaxf(eta, 0.0f)));
        const float s = rcp_approx(ope);                                     // sigmoid(|eta|)
        dl = ((eta >= 0.0f) ? s : (1.0f - s)) - y;
    } else if (LIK == LIK_POISSON) {
        const float r = ex2_approx(eta * 1.4426950408889634f);
        loss = __fmaf_rn(-y, eta, r);
        dl = r - y;
    } else {
        const float res = eta - y;
        loss = 0.5f * res * res;
        dl = res;
    }
}

}  // namespace b2
