"""Host-side convergence diagnostics used for the headline metric (min-ESS/sec, R-hat).

Same definitions as numpyro/diagnostics.py (autocorrelation :101-142, autocovariance :145-155,
effective_sample_size :158-203, gelman_rubin :45-61, split_gelman_rubin :64-80); the reference runs
these in NumPy on the host after sampling, and so does this module.  Inputs are ``[chains, draws,
...]`` arrays; everything is float64.
"""
from __future__ import annotations

import numpy as np

__all__ = ["autocorrelation", "autocovariance", "effective_sample_size", "gelman_rubin",
           "split_gelman_rubin", "summary"]


def _fast_len(n: int) -> int:
    """Smallest 5-smooth integer >= n (FFT-friendly length)."""
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0 and m > 1:
                m //= p
        if m <= 1:
            return n
        n += 1


def _raw_autocov_sums(x: np.ndarray) -> np.ndarray:
    """sum_t xc[t] * xc[t + lag] along the last axis via the Wiener-Khinchin theorem."""
    n = x.shape[-1]
    size = 2 * _fast_len(n)
    xc = x - x.mean(axis=-1, keepdims=True)
    spec = np.fft.rfft(xc, n=size, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    return np.fft.irfft(power, n=size, axis=-1)[..., :n]


def autocorrelation(x, axis: int = 0, bias: bool = True) -> np.ndarray:
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    ac = _raw_autocov_sums(x)
    if not bias:
        ac = ac / np.arange(n, 0.0, -1.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        ac = ac / ac[..., :1]
    return np.moveaxis(ac, -1, axis)


def autocovariance(x, axis: int = 0, bias: bool = True) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return autocorrelation(x, axis, bias) * x.var(axis=axis, keepdims=True)


def _variance_stats(x: np.ndarray):
    n_chains, n_draws = x.shape[:2]
    within = x.var(axis=1, ddof=1).mean(axis=0)
    pooled = within * (n_draws - 1) / n_draws
    if n_chains > 1:
        pooled = pooled + x.mean(axis=1).var(axis=0, ddof=1)
    else:
        within = pooled
    return within, pooled


def gelman_rubin(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    if x.ndim < 2 or x.shape[0] < 2 or x.shape[1] < 2:
        raise ValueError("gelman_rubin needs at least 2 chains of at least 2 draws")
    within, pooled = _variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sqrt(pooled / within)


def split_gelman_rubin(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    if x.ndim < 2 or x.shape[1] < 4:
        raise ValueError("split_gelman_rubin needs at least 4 draws per chain")
    half = x.shape[1] // 2
    return gelman_rubin(np.concatenate([x[:, :half], x[:, -half:]], axis=0))


def effective_sample_size(x, bias: bool = True) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    if x.ndim < 2 or x.shape[1] < 2:
        raise ValueError("effective_sample_size needs [chains, draws, ...] with at least 2 draws")
    n_chains, n_draws = x.shape[:2]
    gamma = autocovariance(x, axis=1, bias=bias)
    within, pooled = _variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        rho = 1.0 - (within - gamma.mean(axis=0)) / pooled
    rho[0] = 1.0
    pairs = rho[:-1:2] + rho[1::2]
    later = np.minimum.accumulate(np.clip(pairs[1:], 0.0, None), axis=0)
    tau = 2.0 * (pairs[0] + later.sum(axis=0)) - 1.0
    return n_chains * n_draws / tau


def summary(samples: dict, group_by_chain: bool = True) -> dict:
    """mean / std / n_eff / r_hat per site (subset of numpyro.diagnostics.summary)."""
    out = {}
    for name, v in samples.items():
        v = np.asarray(v, dtype=np.float64)
        if not group_by_chain:
            v = v[None]
        flat = v.reshape((-1,) + v.shape[2:])
        out[name] = {
            "mean": flat.mean(axis=0), "std": flat.std(axis=0, ddof=1),
            "n_eff": effective_sample_size(v),
            "r_hat": split_gelman_rubin(v) if v.shape[1] >= 4 else np.full(v.shape[2:], np.nan),
        }
    return out
