// Potentials of the registered model families, fused for the engine.
//
// Replaces jax.value_and_grad(potential_fn) (numpyro/infer/hmc_util.py:242-252) over
// potential_energy (infer/util.py:333-358) for: diagonal Gaussian (test target), non-centred
// eight schools (README.md:100-110), and the GLM family -- plain / hierarchical / horseshoe --
// with Bernoulli-logit, Poisson-log or Normal likelihood (examples/covtype.py:66-71,
// examples/horseshoe_regression.py:37-78).  Log-prob formulas: distributions/continuous.py
// Normal :2975-2989, Cauchy :390-406, HalfCauchy :1148-1150, Exponential :715-726, Gamma
// :813-831; discrete.py :263 (-> distributions/util.py:317-320), :1361-1388; the exp-transform
// Jacobian of infer/util.py:325-329 / transforms.py:641-646.  Gradients are hand-derived
// (checked against the fp64 oracle to rtol 1e-5 and by finite differences in tests/).
//
// A GLM evaluation is split in three so that the O(N*D) part can be executed by whichever kernel
// fits the problem (in-warp loop, HBM-streaming pass, tcgen05 GEMM):
//   glm_coef   : beta_j = s_j(z) * u_j                              O(D), per chain
//   likelihood : eta = X beta;  nll = sum_n l(eta_n, y_n);  gbeta = X^T dl/deta   <- the hot op
//   glm_finish : priors, Jacobians and the chain rule back to z     O(D), per chain
#pragma once
#include "common.cuh"

namespace b2 {

enum Family { FAM_DIAG_GAUSSIAN = 0, FAM_EIGHT_SCHOOLS = 1, FAM_GLM = 2 };
enum Likelihood { LIK_BERNOULLI = 0, LIK_POISSON = 1, LIK_NORMAL = 2 };
enum ScalePrior { SCALE_NONE = 0, SCALE_HALFCAUCHY = 1, SCALE_EXPONENTIAL = 2 };

struct FamilySpec {
    int32_t family;
    int32_t D;                 // latent dimension (flat, sorted-site order)
    // data
    long long N; int32_t Dx;   // X [N, Dx] row-major, y [N]
    const float* X; const float* y;
    const float* aux0; const float* aux1;   // gaussian: mu, sigma; eight schools: sigma_j, y_j
    float nll_const;           // poisson: sum_n lgamma(y_n + 1), precomputed at create()
    // GLM structure; offsets into z (-1 = site absent).  Sorted-site layouts:
    //   plain: coefs | horseshoe: lambdas, (prec_obs), tau, unscaled_betas | hierarchical: coefs, tau
    int32_t likelihood, off_lambda, off_tau, off_prec, off_u, gscale, g0, g1;
    float tau_scale, mu_scale;
    // Energy-conserving subsampling (HMCECS, numpyro/infer/hmc_gibbs.py:502-690; plain GLM only): the likelihood is estimated
    // from ecs_m of the N rows with a Taylor-proxy control variate around `ecs_ref` (contrib/ecs_proxies.py:23-50, 95-300)
    int32_t ecs_m;                 // 0: off
    int32_t ecs_degree;            // Taylor proxy degree 1 / 2; 0 = no proxy (plate-scaled subsample likelihood)
    const int32_t* ecs_idx;        // [ecs_m] subsample rows of THIS chain (the caller offsets the per-chain table)
    const float* ecs_eta_ref;      // [N]  x_i . ref
    const float* ecs_ref;          // [Dx]
    const float* ecs_G;            // [Dx] gradient of the full-data log-likelihood at ref
    const float* ecs_H;            // [Dx][Dx] its Hessian (degree 2)
    float ecs_L0;                  // full-data log-likelihood at ref
    // Conditioning on Gibbs sites (HMCGibbs, numpyro/infer/hmc_gibbs.py:38-192): the chain's vector holds only the cond_Dfree
    // free coordinates; the potential is the full model's with the fixed coordinates substituted (warp regime)
    int32_t cond_Dfree;            // 0: off
    const int32_t* cond_map;       // [D] index into the chain's reduced vector, or -1 for a fixed (Gibbs) coordinate
    const float* cond_val;         // [D + 1] of THIS chain: unconstrained values of the fixed coordinates; [D] = what to add to U (the
                                   // exp-transform Jacobians of fixed positive sites, absent from the conditioned model)
    float* cond_scratch;           // [2 D] of THIS chain
};

constexpr float kLogSqrt2Pi = 0.918938533204672742f;
constexpr float kLog2 = 0.693147180559945309f;
constexpr float kLogPi = 1.144729885849400174f;

// (lane_sync(): common.cuh)

// ---- per-observation loss: value and d/d eta of the negative log-likelihood ------------------
B2_HD void glm_loss(int lik, float eta, float y, float& loss, float& dl) {
    if (lik == LIK_BERNOULLI) {          // binary_cross_entropy_with_logits (distributions/util.py:317-320)
        const float e = expf(-fabsf(eta));
        loss = fmaxf(eta, 0.0f) + log1pf(e) - eta * y;
        const float s = 1.0f / (1.0f + e);                 // sigmoid(|eta|)
        dl = ((eta >= 0.0f) ? s : (1.0f - s)) - y;         // sigmoid(eta) - y
    } else if (lik == LIK_POISSON) {     // Poisson.log_prob (discrete.py:1388) with rate = exp(eta)
        const float r = expf(eta);
        loss = r - y * eta;                               // + lgamma(y+1): FamilySpec.nll_const
        dl = r - y;
    } else {                             // Normal: 0.5 * res^2 (precision applied in glm_finish)
        const float res = eta - y;
        loss = 0.5f * res * res;
        dl = res;
    }
}

#if defined(__CUDACC__)
// Streaming-kernel variant: hardware ex2/lg2/rcp approximations (relative error ~1e-7, far inside
// the rtol 1e-5 parity budget once summed over rows); no libm slow paths in the hot loop.
__device__ __forceinline__ void glm_loss_fast(int lik, float eta, float y, float& loss, float& dl) {
    if (lik == LIK_BERNOULLI) {
        const float e = __expf(-fabsf(eta));
        const float ope = 1.0f + e;
        loss = (fmaxf(eta, 0.0f) - eta * y) + __logf(ope);
        const float s = __fdividef(1.0f, ope);
        dl = ((eta >= 0.0f) ? s : (1.0f - s)) - y;
    } else if (lik == LIK_POISSON) {
        const float r = __expf(eta);
        loss = r - y * eta;
        dl = r - y;
    } else {
        const float res = eta - y;
        loss = 0.5f * res * res;
        dl = res;
    }
}
#endif

B2_HD float glm_scale_at(const FamilySpec& f, const float* z, int j) {
    float s = 1.0f;
    if (f.off_lambda >= 0) s = s * expf(z[f.off_lambda + j]);
    if (f.gscale != SCALE_NONE && j >= f.g0 && j < f.g1) s = s * expf(z[f.off_tau]);
    return s;
}

B2_HD void glm_coef(const FamilySpec& f, const float* z, float* beta) {
    B2_FOR_D(j, f.Dx) beta[j] = glm_scale_at(f, z, j) * z[f.off_u + j];
}

// nll / gbeta are the raw likelihood sums (for LIK_NORMAL: 0.5*sum res^2 and X^T res).
B2_HD void glm_finish(const FamilySpec& f, const float* z, float nll, const float* gbeta, float& u_out, float* g) {
    const int Dx = f.Dx;
#if defined(__CUDA_ARCH__)
    if (f.off_lambda < 0 && f.gscale == SCALE_NONE && f.likelihood != LIK_NORMAL && Dx <= 64) {
        // plain GLM (coefs ~ N(0, 1)) with at most two coefficients per lane: loads first, then the same arithmetic as below
        // (lane partial in element order + butterfly; scale = 1, prec = 1 multiplications kept so that the bits agree)
        const int lane = (int)(threadIdx.x & 31u);
        const int d0 = lane, d1 = lane + 32;
        const bool a0 = d0 < Dx, a1 = d1 < Dx;
        float z0 = 0.0f, z1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
        if (a0) { z0 = z[f.off_u + d0]; b0 = gbeta[d0]; }
        if (a1) { z1 = z[f.off_u + d1]; b1 = gbeta[d1]; }
        float acc = 0.0f;
        if (a0) acc = acc + (0.5f * z0) * z0;
        if (a1) acc = acc + (0.5f * z1) * z1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = acc + __shfl_xor_sync(0xFFFFFFFFu, acc, off);
        float U = acc + (float)Dx * kLogSqrt2Pi;
        if (a0) g[f.off_u + d0] = z0 + 1.0f * (1.0f * b0);
        if (a1) g[f.off_u + d1] = z1 + 1.0f * (1.0f * b1);
        U = U + (nll + f.nll_const);
        u_out = U;
        return;
    }
#endif
    float prec = 1.0f;
    if (f.likelihood == LIK_NORMAL) prec = expf(z[f.off_prec]);
    // (wide models -- the GEMM regime's few-but-wide chains: four elements of a lane in flight per batch, see common.cuh; the
    //  per-element arithmetic and the order of the additions are those of the plain loops: identical bits)
    const bool has_l = f.off_lambda >= 0, has_t = f.gscale != SCALE_NONE;
    const float zt_ = has_t ? z[f.off_tau] : 0.0f;
    auto scale_of = [&](float zl, int j) {           // glm_scale_at with the operands already loaded
        float s = 1.0f;
        if (has_l) s = s * expf(zl);
        if (has_t && j >= f.g0 && j < f.g1) s = s * expf(zt_);
        return s;
    };
    // coefficient block: u ~ N(0, 1)
    B2_LAPQ(-1);
    float acc = lane_sum_wide(Dx, [&](int j) { const float u = z[f.off_u + j]; return 0.5f * u * u; });
    float U = acc + (float)Dx * kLogSqrt2Pi;
    B2_LAPQ(8);
    for_d_wide(Dx, [&](int j) { Vals v; v.x[0] = z[f.off_u + j]; v.x[1] = has_l ? z[f.off_lambda + j] : 0.0f; v.x[2] = gbeta[j]; return v; },
               [&](int j, const Vals& v) { g[f.off_u + j] = v.x[0] + scale_of(v.x[1], j) * (prec * v.x[2]); });
    B2_LAPQ(9);
    if (has_l) {                         // lambdas ~ HalfCauchy(1), z = log lambda
        float a = lane_sum_wide(Dx, [&](int j) {
            const float zl = z[f.off_lambda + j];
            return log1pf(expf(2.0f * zl)) - zl;
        });
        U = U + a + (float)Dx * (kLogPi - kLog2);
        for_d_wide(Dx, [&](int j) { Vals v; v.x[0] = z[f.off_lambda + j]; v.x[1] = z[f.off_u + j]; v.x[2] = gbeta[j]; return v; },
                   [&](int j, const Vals& v) {
            const float zl = v.x[0];
            const float l2 = expf(2.0f * zl);
            const float beta = scale_of(zl, j) * v.x[1];
            // d/dz of [log1p(l2) - zl] = 2*l2/(1+l2) - 1; inf/inf guarded
            const float frac = is_inf(l2) ? 1.0f : (l2 / (1.0f + l2));
            g[f.off_lambda + j] = beta * (prec * v.x[2]) + 2.0f * frac - 1.0f;
        });
    }
    if (f.gscale != SCALE_NONE) {        // global scale tau, z = log tau
        const float zt = z[f.off_tau];
        const float tau = expf(zt);
        const float dot = lane_sum_wide(Dx, [&](int j) {
            if (j < f.g0 || j >= f.g1) return 0.0f;
            return (glm_scale_at(f, z, j) * z[f.off_u + j]) * (prec * gbeta[j]);
        });
        float pr, dpr;
        if (f.gscale == SCALE_HALFCAUCHY) {
            const float q = (tau / f.tau_scale) * (tau / f.tau_scale);
            pr = kLogPi - kLog2 + logf(f.tau_scale) + log1pf(q) - zt;
            const float frac = is_inf(q) ? 1.0f : (q / (1.0f + q));
            dpr = 2.0f * frac - 1.0f;
        } else {
            const float rate = 1.0f / f.tau_scale;
            pr = rate * tau - logf(rate) - zt;
            dpr = rate * tau - 1.0f;
        }
        U = U + pr;
        if (lane_first() == 0) g[f.off_tau] = dot + dpr;
    }
    if (f.likelihood == LIK_NORMAL) {    // sigma = prec^-1/2, prec ~ Gamma(3, 1), z = log prec
        const float zp = z[f.off_prec];
        const float Nf = (float)f.N;
        U = U + prec * nll - 0.5f * Nf * zp + Nf * kLogSqrt2Pi;
        U = U + (prec - 3.0f * zp + 0.693147180559945309f);          // -[2 zp - prec - lgamma(3)] - zp
        if (lane_first() == 0) g[f.off_prec] = prec * nll - 0.5f * Nf + prec - 3.0f;
    } else {
        U = U + (nll + f.nll_const);
    }
    u_out = U;
    B2_LAPQ(10);
}

// ---- HMCECS potential (one warp): prior + bias-corrected subsample estimate of the log-likelihood -----------------------
// perturbed_method (contrib/ecs_proxies.py:23-50) with the Taylor proxy (:95-300) written out for a GLM: per subsampled row
// only eta_i = x_i . z and e0_i = x_i . ref are needed, because the row's log-likelihood depends on z through eta_i alone:
//   proxy_i(z) = l(e0) + l'(e0) a + 0.5 l''(e0) a^2,  a = eta - e0   (= ll_i(ref) + g_i . dz + 0.5 dz' H_i dz)
//   diff_i = l(eta) - proxy_i;   ll = proxy_all + N mean(diff) - 0.5 (N^2 / m) var(diff)
// scratch: >= 2 m + Dx floats.
B2_HD void potential_ecs_inwarp(const FamilySpec& f, const float* z, float* scratch, float& u_out, float* g) {
    const int m = f.ecs_m, Dx = f.Dx, deg = f.ecs_degree;
    float* diff = scratch; float* wc = scratch + m; float* dz = scratch + 2 * m;
    const float Nf = (float)f.N, mf = (float)m;
    if (deg > 0) { B2_FOR_D(j, Dx) dz[j] = z[f.off_u + j] - f.ecs_ref[j]; }
    lane_sync();
    // pass 1 over the subsample: diff_k and c_k = d diff_k / d eta
    const float sum_diff = lane_sum(m, [&](int k) {
        const long long i = f.ecs_idx[k];
        const float* x = f.X + (size_t)i * Dx;
        float eta = 0.0f;
        for (int j = 0; j < Dx; ++j) eta = fmaf(x[j], z[f.off_u + j], eta);
        const float y = f.y[i];
        float loss, dl;
        glm_loss(f.likelihood, eta, y, loss, dl);
        float ll = -loss;
        if (f.likelihood == LIK_POISSON) ll = ll - lgammaf(y + 1.0f);
        float d = ll, c = -dl;
        if (deg > 0) {
            const float e0 = f.ecs_eta_ref[i];
            float loss0, dl0;
            glm_loss(f.likelihood, e0, y, loss0, dl0);
            float ll0 = -loss0;
            if (f.likelihood == LIK_POISSON) ll0 = ll0 - lgammaf(y + 1.0f);
            const float a = eta - e0;
            float proxy = ll0 + (-dl0) * a;
            c = c - (-dl0);
            if (deg == 2) {
                float l2;                                     // l''(e0)
                if (f.likelihood == LIK_BERNOULLI) { const float s = 1.0f / (1.0f + expf(-e0)); l2 = -(s * (1.0f - s)); }
                else l2 = -expf(e0);
                proxy = proxy + 0.5f * l2 * a * a;
                c = c - l2 * a;
            }
            d = ll - proxy;
        }
        diff[k] = d; wc[k] = c;
        return d;
    });
    lane_sync();
    const float mean = sum_diff / mf;
    float ll_est, w_var = 0.0f;
    if (deg > 0) {
        const float ss = lane_sum(m, [&](int k) { const float t = diff[k] - mean; return t * t; });
        const float var = ss / mf;                            // jnp.var: population variance
        // proxy over all rows: L0 + G . dz + 0.5 dz' H dz
        float lin, quad = 0.0f;
        lin = lane_sum(Dx, [&](int j) { return f.ecs_G[j] * dz[j]; });
        if (deg == 2) quad = lane_sum(Dx, [&](int i) {
            const float* row = f.ecs_H + (size_t)i * Dx;
            float a = 0.0f;
            for (int j = 0; j < Dx; ++j) a = fmaf(row[j], dz[j], a);
            return a * dz[i];
        });
        ll_est = (f.ecs_L0 + lin + 0.5f * quad) + Nf * mean - 0.5f * ((Nf * Nf / mf) * var);
        w_var = Nf * Nf / (mf * mf);
    } else {
        ll_est = (Nf / mf) * sum_diff;                        // plate scaling N / m (primitives.py plate, no proxy)
    }
    // weights of the rows in the gradient: d ll / d diff_k = N/m - (N^2/m^2)(diff_k - mean)
    B2_FOR_D(k, m) wc[k] = wc[k] * (Nf / mf - w_var * (diff[k] - mean));
    lane_sync();
    const float prior = lane_sum(Dx, [&](int j) { const float u = z[f.off_u + j]; return 0.5f * u * u; });
    B2_FOR_D(j, Dx) {
        float a = 0.0f;
        for (int k = 0; k < m; ++k) a = fmaf(f.X[(size_t)f.ecs_idx[k] * Dx + j], wc[k], a);
        if (deg > 0) {
            a = a + f.ecs_G[j];
            if (deg == 2) { const float* row = f.ecs_H + (size_t)j * Dx; float h = 0.0f; for (int q = 0; q < Dx; ++q) h = fmaf(row[q], dz[q], h); a = a + h; }
        }
        g[f.off_u + j] = z[f.off_u + j] - a;
    }
    u_out = (prior + (float)Dx * kLogSqrt2Pi) - ll_est;
    lane_sync();
}

// ---- whole potential inside one warp (tiny models, regime R1) ---------------------------------
// scratch: >= N + Dx floats private to the chain (residuals, beta), gtmp: D floats.
B2_HD void potential_inwarp_full(const FamilySpec& f, const float* z, float* scratch, float& u_out, float* g) {
    if (f.family == FAM_DIAG_GAUSSIAN) {
        const float* mu = f.aux0; const float* sg = f.aux1;
        u_out = 0.5f * lane_sum(f.D, [&](int d) { const float t = (z[d] - mu[d]) / sg[d]; return t * t; });
        B2_FOR_D(d, f.D) g[d] = ((z[d] - mu[d]) / sg[d]) / sg[d];
        return;
    }
    if (f.family == FAM_EIGHT_SCHOOLS) {  // z = [mu, log tau, theta_base[J]]
        const int J = f.D - 2;
        const float* sg = f.aux0; const float* y = f.aux1;
        const float mu = z[0], zt = z[1];
        const float tau = expf(zt);
        float s_res2, s_d;                 // sum 0.5*res^2 + log sigma ; sum dtheta
        lane_sum2(J, [&](int j, float& a, float& b) {
            const float res = (y[j] - (mu + tau * z[2 + j])) / sg[j];
            a = 0.5f * res * res + logf(sg[j]);
            b = res / sg[j];
        }, s_res2, s_d);
        float s_tb2, s_dtb;
        lane_sum2(J, [&](int j, float& a, float& b) {
            const float tb = z[2 + j];
            const float res = (y[j] - (mu + tau * tb)) / sg[j];
            a = 0.5f * tb * tb;
            b = (res / sg[j]) * tb;
        }, s_tb2, s_dtb);
        const float ms = f.mu_scale, ts = f.tau_scale;
        const float q = (tau / ts) * (tau / ts);
        float U = 0.5f * (mu / ms) * (mu / ms) + logf(ms) + kLogSqrt2Pi;
        U = U + (kLogPi - kLog2 + logf(ts) + log1pf(q) - zt);
        U = U + s_tb2 + s_res2 + 2.0f * (float)J * kLogSqrt2Pi;
        u_out = U;
        const float frac = is_inf(q) ? 1.0f : (q / (1.0f + q));
        if (lane_first() == 0) {
            g[0] = mu / (ms * ms) - s_d;
            g[1] = 2.0f * frac - 1.0f - s_dtb * tau;
        }
        B2_FOR_D(j, J) {
            const float tb = z[2 + j];
            const float res = (y[j] - (mu + tau * tb)) / sg[j];
            g[2 + j] = tb - (res / sg[j]) * tau;
        }
        return;
    }
    if (f.ecs_m > 0) { potential_ecs_inwarp(f, z, scratch, u_out, g); return; }
    // FAM_GLM, small N: lanes over rows for eta / residuals, then lanes over columns for X^T r
    float* beta = scratch; float* resid = scratch + f.Dx;
    glm_coef(f, z, beta);
    lane_sync();
    const int N = (int)f.N, Dx = f.Dx;
    const float nll = lane_sum(N, [&](int n) {
        const float* x = f.X + (size_t)n * Dx;
        float eta = 0.0f;
        for (int j = 0; j < Dx; ++j) eta = fmaf(x[j], beta[j], eta);
        float loss, dl;
        glm_loss(f.likelihood, eta, f.y[n], loss, dl);
        resid[n] = dl;
        return loss;
    });
    lane_sync();
    B2_FOR_D(j, Dx) {
        float a = 0.0f;
        for (int n = 0; n < N; ++n) a = fmaf(f.X[(size_t)n * Dx + j], resid[n], a);
        beta[j] = a;                          // reuse as gbeta (own lane's slots only)
    }
    lane_sync();
    glm_finish(f, z, nll, beta, u_out, g);
}

// The potential the chain's state machine sees: the family's, or -- conditioned on Gibbs sites (HMCGibbs) -- the family's with the
// fixed coordinates substituted: HMC sites -> full vector, full potential, gradient of the free coordinates back.
B2_HD void potential_inwarp(const FamilySpec& f, const float* z, float* scratch, float& u_out, float* g) {
    if (f.cond_Dfree <= 0) { potential_inwarp_full(f, z, scratch, u_out, g); return; }
    float* zf = f.cond_scratch; float* gf = f.cond_scratch + f.D;
    B2_FOR_D(d, f.D) { const int k = f.cond_map[d]; zf[d] = (k >= 0) ? z[k] : f.cond_val[d]; }
    lane_sync();
    potential_inwarp_full(f, zf, scratch, u_out, gf);
    lane_sync();
    B2_FOR_D(d, f.D) { const int k = f.cond_map[d]; if (k >= 0) g[k] = gf[d]; }
    u_out = u_out + f.cond_val[f.D];
    lane_sync();
}

}  // namespace b2
