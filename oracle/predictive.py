"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the consumers of the collected samples,
numpyro/infer/util.py ``log_likelihood`` (:1133-1188) and ``Predictive`` / ``_predictive`` (:838-1131), for the
registered families.  fp64 arithmetic rounded once; PRNG draws through oracle/prng.py.

Key plumbing of ``_predictive`` (util.py:916-920 + handlers.py ``seed.process_message`` :887-897): the call's key is split
into one key per posterior sample when there is more than one sample (the key itself otherwise); inside one prediction
every latent site is substituted, so the observed site is the only message without a value and receives
``split(sample_key)[1]``.  Draws: ``Normal.sample`` = loc + scale * normal(key, shape) (continuous.py:2961-2967),
``BernoulliLogits.sample`` = uniform(key, shape) < expit(logits) (discrete.py:226-241, returned as int32 0/1).
"""
from __future__ import annotations

import numpy as np

from . import prng
from .families import GLM, EightSchools, LOG_SQRT_2PI

F = np.float32


def sample_keys(rng_key, num_samples: int) -> np.ndarray:
    """util.py:916-918: ``random.split(rng_key, num_samples)`` if num_samples > 1 else the key itself."""
    rng_key = np.asarray(rng_key, np.uint32).reshape(2)
    return rng_key[None].copy() if num_samples <= 1 else prng.split(rng_key, num_samples)


def obs_key(sample_key) -> np.ndarray:
    """handlers.py:896: ``self.rng_key, rng_key_sample = random.split(self.rng_key)`` at the first un-valued sample site."""
    return prng.split(np.asarray(sample_key, np.uint32).reshape(2))[1]


def _eta_glm(fam: GLM, z64):
    p = fam._split(z64)
    return fam.X @ (fam.coef_scale(p) * p[fam.coef_name]), p


def obs_location(fam, z) -> np.ndarray:
    """Mean parameter of the observed site (fp64): theta for eight schools, the linear predictor eta for a GLM."""
    z64 = np.asarray(z, np.float64)
    if isinstance(fam, EightSchools):
        return z64[0] + np.exp(z64[1]) * z64[2:]
    return _eta_glm(fam, z64)[0]


def log_likelihood(fam, z) -> np.ndarray:
    """``site['fn'].log_prob(site['value'])`` of the observed site for one unconstrained sample ``z`` (util.py:1158-1166)."""
    z64 = np.asarray(z, np.float64)
    with np.errstate(all="ignore"):
        if isinstance(fam, EightSchools):
            r = (fam.y - obs_location(fam, z64)) / fam.sigma
            return (-0.5 * r * r - np.log(fam.sigma) - LOG_SQRT_2PI).astype(F)      # continuous.py:2975-2989
        eta, p = _eta_glm(fam, z64)
        if fam.likelihood == "bernoulli":                                        # discrete.py:263 (BernoulliLogits.log_prob)
            return (-(np.maximum(eta, 0) + np.log1p(np.exp(-np.abs(eta))) - eta * fam.y)).astype(F)
        if fam.likelihood == "poisson":                                          # discrete.py:1388
            return (fam.y * eta - np.exp(eta) - fam._lgam).astype(F)
        zp = p["prec_obs"][0]
        res = fam.y - eta
        return (-0.5 * np.exp(zp) * res * res + 0.5 * zp - LOG_SQRT_2PI).astype(F)


def predictive(fam, z, sample_key) -> np.ndarray:
    """One posterior-predictive draw of the observed site for the unconstrained sample ``z`` with the sample's key."""
    k = obs_key(sample_key)
    loc = obs_location(fam, z)
    n = loc.shape[0]
    with np.errstate(all="ignore"):
        if isinstance(fam, EightSchools):
            return (loc + fam.sigma * prng.normal(k, n).astype(np.float64)).astype(F)
        if fam.likelihood == "bernoulli":
            return (prng.uniform(k, n) < (1.0 / (1.0 + np.exp(-loc))).astype(F)).astype(F)
        if fam.likelihood == "normal":
            zp = np.asarray(z, np.float64)[dict((nm, off) for nm, off, _ in fam.layout)["prec_obs"]]
            return (loc + np.exp(-0.5 * zp) * prng.normal(k, n).astype(np.float64)).astype(F)
    raise NotImplementedError("Poisson predictive draws are not restated")


def bernoulli_margin(fam, z, sample_key) -> np.ndarray:
    """|u - p| per observation: a draw may legitimately differ from the oracle's only where this is at rounding level."""
    k = obs_key(sample_key)
    loc = obs_location(fam, z)
    return np.abs(prng.uniform(k, loc.shape[0]).astype(np.float64) - 1.0 / (1.0 + np.exp(-loc)))
