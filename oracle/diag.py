"""Oracle: convergence diagnostics (TEST INFRASTRUCTURE ONLY).

Independent restatement of numpyro/diagnostics.py: _compute_chain_variance_stats :29-42,
gelman_rubin :45-61, split_gelman_rubin :64-80, autocorrelation :101-142, autocovariance
:145-155, effective_sample_size :158-203.  The reference uses an FFT; this oracle computes the
autocovariance by the direct O(S^2) sum so that it cross-checks the product's FFT version.
Pins: tests/test_oracle_diag.py restates test/test_diagnostics.py:60-108.
"""
from __future__ import annotations

import numpy as np


def autocovariance(x, axis=0):
    """Biased (1/N) autocovariance along ``axis`` (diagnostics.py:145-155)."""
    x = np.moveaxis(np.asarray(x, np.float64), axis, -1)
    n = x.shape[-1]
    xc = x - x.mean(axis=-1, keepdims=True)
    out = np.empty_like(xc)
    for lag in range(n):
        out[..., lag] = np.sum(xc[..., : n - lag] * xc[..., lag:], axis=-1) / n
    return np.moveaxis(out, -1, axis)


def autocorrelation(x, axis=0, bias=True):
    """diagnostics.py:101-142: autocovariance normalised by lag 0 (``bias=False`` rescales by N/(N-lag))."""
    x = np.moveaxis(np.asarray(x, np.float64), axis, -1)
    n = x.shape[-1]
    ac = np.moveaxis(autocovariance(x, axis=-1), -1, -1)
    if not bias:
        ac = ac * n / np.arange(n, 0.0, -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ac = ac / ac[..., :1]
    return np.moveaxis(ac, -1, axis)


def _chain_variance_stats(x):
    """Within-chain variance W and the pooled estimator V (diagnostics.py:29-42).
    With a single chain the reference returns W := V = W*(N-1)/N."""
    c, n = x.shape[0], x.shape[1]
    w = x.var(axis=1, ddof=1).mean(axis=0)
    v = w * (n - 1) / n
    if c > 1:
        v = v + x.mean(axis=1).var(axis=0, ddof=1)
    else:
        w = v
    return w, v


def gelman_rubin(x):
    x = np.asarray(x, np.float64)
    assert x.ndim >= 2 and x.shape[0] >= 2 and x.shape[1] >= 2
    w, v = _chain_variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sqrt(v / w)


def split_gelman_rubin(x):
    x = np.asarray(x, np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 4
    h = x.shape[1] // 2
    return gelman_rubin(np.concatenate([x[:, :h], x[:, -h:]], axis=0))


def effective_sample_size(x, bias=True):
    """Geyer initial-monotone-sequence ESS as in Stan (diagnostics.py:158-203)."""
    x = np.asarray(x, np.float64)
    assert x.ndim >= 2 and x.shape[1] >= 2
    c, n = x.shape[0], x.shape[1]
    gamma = autocovariance(x, axis=1)                     # biased (1/N), per chain
    if not bias:
        shape = [1] * x.ndim
        shape[1] = n
        gamma = gamma * (n / np.arange(n, 0.0, -1)).reshape(shape)
    w, v = _chain_variance_stats(x)
    with np.errstate(invalid="ignore", divide="ignore"):
        rho = 1.0 - (w - gamma.mean(axis=0)) / v
    rho[0] = 1.0
    m = n // 2 if n % 2 == 0 else (n - 1) // 2
    pairs = rho[0:2 * m:2] + rho[1:2 * m + 1:2][:m]
    # first pair kept as is; later pairs clipped at 0, then running minimum among themselves
    tail = np.minimum.accumulate(np.clip(pairs[1:], 0.0, None), axis=0)
    tau = -1.0 + 2.0 * (pairs[:1].sum(axis=0) + tail.sum(axis=0))
    return c * n / tau
