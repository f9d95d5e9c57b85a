// The step AFTER the hot path: per-observation log-likelihoods and posterior-predictive draws
// over a batch of collected samples, numpyro/infer/util.py log_likelihood :1133-1188 and Predictive / _predictive :838-1131.
// For the GLM families both are a [S samples x N rows] product X beta_s with an elementwise epilogue:
//   log_likelihood : site["fn"].log_prob(y)          (discrete.py:263, :1388; continuous.py:2975-2989)
//   predictive     : site["fn"].sample(key_s)        (Bernoulli: uniform(key, (N,)) < expit(eta), discrete.py:226-241;
//                                                     Normal: loc + scale * normal(key, (N,)), continuous.py:2961-2967)
// key_s is the key the seed handler hands to the observed site for sample s (numpyro_b200/predictive.py derives it).
#pragma once
#include "common.cuh"
#include "families.cuh"
#include "prng.cuh"

namespace b2 {

constexpr int kPrRows = 128, kPrSamples = 8, kPrCols = 32;

// MODE 0: log-likelihood, 1: predictive draw.  z [S][D] unconstrained samples, out [S][N].
template <int MODE>
__global__ void __launch_bounds__(kPrRows) k_glm_rows(FamilySpec f, const float* __restrict__ z, const uint32_t* __restrict__ keys,
                                                       long long S, float* __restrict__ out) {
    __shared__ float xs[kPrRows][kPrCols + 1];
    __shared__ float bs[kPrSamples][kPrCols];
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kPrRows, s0 = (long long)blockIdx.y * kPrSamples;
    const long long row = row0 + tid;
    float acc[kPrSamples];
#pragma unroll
    for (int s = 0; s < kPrSamples; ++s) acc[s] = 0.0f;
    for (int c0 = 0; c0 < f.Dx; c0 += kPrCols) {
        for (int i = tid; i < kPrRows * kPrCols; i += kPrRows) {            // X tile, coalesced along the columns
            const int r = i / kPrCols, k = i - r * kPrCols;
            xs[r][k] = (row0 + r < f.N && c0 + k < f.Dx) ? f.X[(row0 + r) * f.Dx + c0 + k] : 0.0f;
        }
        for (int i = tid; i < kPrSamples * kPrCols; i += kPrRows) {         // betas of the samples of this block
            const int s = i / kPrCols, k = i - s * kPrCols;
            float b = 0.0f;
            if (s0 + s < S && c0 + k < f.Dx) { const float* zs = z + (s0 + s) * f.D; b = glm_scale_at(f, zs, c0 + k) * zs[f.off_u + c0 + k]; }
            bs[s][k] = b;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kPrCols; ++k) {
            const float x = xs[tid][k];
#pragma unroll
            for (int s = 0; s < kPrSamples; ++s) acc[s] = fmaf(x, bs[s][k], acc[s]);
        }
        __syncthreads();
    }
    if (row >= f.N) return;
    const float y = f.y[row];
    for (int s = 0; s < kPrSamples && s0 + s < S; ++s) {
        const float eta = acc[s];
        const float* zs = z + (s0 + s) * f.D;
        float v;
        if (MODE == 0) {
            if (f.likelihood == LIK_BERNOULLI) v = -((fmaxf(eta, 0.0f) + log1pf(expf(-fabsf(eta)))) - eta * y);
            else if (f.likelihood == LIK_POISSON) v = y * eta - expf(eta) - lgammaf(y + 1.0f);
            else { const float zp = zs[f.off_prec], res = y - eta; v = -0.5f * expf(zp) * res * res + 0.5f * zp - kLogSqrt2Pi; }
        } else {
            Key k; k.a = keys[2 * (s0 + s)]; k.b = keys[2 * (s0 + s) + 1];
            if (f.likelihood == LIK_BERNOULLI) v = (uniform01_at(k, (uint32_t)row) < 1.0f / (1.0f + expf(-eta))) ? 1.0f : 0.0f;
            else v = eta + expf(-0.5f * zs[f.off_prec]) * normal_at(k, (uint32_t)row);      // sigma = prec^-1/2
        }
        out[(s0 + s) * f.N + row] = v;
    }
}

// eight schools (README.md:100-110): obs_j ~ Normal(mu + tau * theta_base_j, sigma_j)
template <int MODE>
__global__ void k_eight_rows(FamilySpec f, const float* __restrict__ z, const uint32_t* __restrict__ keys, long long S, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int J = f.D - 2;
    if (i >= S * J) return;
    const long long s = i / J; const int j = (int)(i - s * J);
    const float* zs = z + s * f.D;
    const float theta = zs[0] + expf(zs[1]) * zs[2 + j], sg = f.aux0[j];
    if (MODE == 0) { const float r = (f.aux1[j] - theta) / sg; out[i] = -0.5f * r * r - logf(sg) - kLogSqrt2Pi; }
    else { Key k; k.a = keys[2 * s]; k.b = keys[2 * s + 1]; out[i] = theta + sg * normal_at(k, (uint32_t)j); }
}

}  // namespace b2
