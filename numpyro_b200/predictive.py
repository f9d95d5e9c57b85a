"""The consumers of the collected samples: ``Predictive`` and ``log_likelihood``.

Mirrors numpyro/infer/util.py ``Predictive`` (:927-1131, ``_predictive`` :838-924) and ``log_likelihood`` (:1133-1188)
for the declared model families: the posterior samples (the constrained dict ``MCMC.get_samples()`` returns) are mapped
back to the flat unconstrained layout and handed to the engine, which computes the ``[samples x observations]`` product
and the observed site's ``log_prob`` / ``sample`` in one kernel (csrc/predict.cuh through b200nuts_log_likelihood /
b200nuts_predict).  Key plumbing follows the reference so that identical keys give identical draws: the call's key is
split once per posterior sample (util.py:916-918) and the observed site -- the only site without a substituted value --
receives ``split(sample_key)[1]`` from the seed handler (handlers.py:887-897).  No CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _capi, families, random as b2random
from .engine import Engine

_CHUNK_FLOATS = 1 << 28          # <= 1 GiB of device output per launch; larger requests are processed in sample chunks


def _batch_shape(posterior_samples: Dict[str, np.ndarray], batch_ndims: int):
    """util.py:1047-1075 / :1169-1179: every site must share the leading ``batch_ndims`` dimensions."""
    shape, first = None, None
    for name, v in posterior_samples.items():
        s = tuple(np.shape(v)[:batch_ndims])
        if shape is not None and s != shape:
            raise ValueError(f"Batch shapes at site {name} and {first} should be the same, but got {s} and {shape}")
        shape, first = s, name
    if shape is None:
        raise NotImplementedError("prior predictive (no posterior_samples) is not on the accelerated path")
    return shape


def _unconstrained(bound, posterior_samples, batch_shape, D) -> np.ndarray:
    """Constrained latent sites -> flat unconstrained ``[n, D]`` (inverse of ExpTransform for positive sites,
    transforms.py:635-646); deterministic sites in ``posterior_samples`` are ignored (``exclude_deterministic``)."""
    n = int(np.prod(batch_shape)) if batch_shape else 1
    z = np.zeros((n, D), np.float32)
    for s in bound.latent_sites:
        if s.name not in posterior_samples:
            raise NotImplementedError(f"posterior_samples must hold every latent site (missing {s.name!r}): sampling "
                                      "missing latent sites from the prior is not on the accelerated path")
        v = np.asarray(posterior_samples[s.name], np.float32).reshape(n, s.size)
        if s.positive:
            with np.errstate(all="ignore"):
                v = np.log(v)
        z[:, s.z_offset:s.z_offset + s.size] = v
    return z


def _engine_for(model, args, kwargs):
    if not isinstance(model, families.Model):
        raise TypeError("model must be a numpyro_b200.families.Model (a declared model family)")
    bound = model.bind(*args, _predict=True, **kwargs)
    cfg = dict(bound.cfg)
    cfg.update(num_chains=1, regime=_capi.REGIME_WARP)      # no tile images: the row kernels read the caller's X directly
    e = Engine(device=torch.device("cuda", torch.cuda.current_device()), X=bound.X, y=bound.y, aux=bound.aux, **cfg)
    return bound, e


def _in_chunks(n, n_obs, fn):
    per = max(1, _CHUNK_FLOATS // max(1, n_obs))
    parts = [fn(lo, min(n, lo + per)).cpu().numpy() for lo in range(0, n, per)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, n_obs), np.float32)


def log_likelihood(model, posterior_samples, *args, parallel=False, batch_ndims=1, **kwargs) -> Dict[str, np.ndarray]:
    """``numpyro.infer.log_likelihood`` (util.py:1133-1188): log-probability of the observed site at every observation,
    for every posterior sample; result ``{obs_site: [*batch_shape, n_obs]}``."""
    shape = _batch_shape(posterior_samples, batch_ndims)
    bound, e = _engine_for(model, args, kwargs)
    try:
        z = torch.from_numpy(_unconstrained(bound, posterior_samples, shape, e.D)).to(e.device)
        n_obs = int(e.lib.b200nuts_obs_count(e.h))
        out = _in_chunks(z.shape[0], n_obs, lambda lo, hi: e.log_likelihood(z[lo:hi]))
    finally:
        e.close()
    return {model.obs_name: out.reshape(tuple(shape) + (n_obs,))}


class Predictive:
    """``numpyro.infer.Predictive`` (util.py:927-1131) for posterior samples of a declared family."""

    def __init__(self, model, posterior_samples: Optional[Dict] = None, *, guide=None, params=None, num_samples=None,
                 return_sites=None, infer_discrete=False, parallel=False, batch_ndims: Optional[int] = None,
                 exclude_deterministic: bool = True):
        if guide is not None or params is not None or infer_discrete:
            raise NotImplementedError("guide / params / infer_discrete are outside the accelerated path")
        if posterior_samples is None:
            raise NotImplementedError("prior predictive (no posterior_samples) is not on the accelerated path")
        if not isinstance(model, families.Model):
            raise TypeError("model must be a numpyro_b200.families.Model (a declared model family)")
        batch_ndims = 1 if batch_ndims is None else int(batch_ndims)          # util.py:1036-1037 (guide is None)
        shape = _batch_shape(posterior_samples, batch_ndims)
        n = int(np.prod(shape)) if shape else 1
        if num_samples is not None and num_samples != n:                     # util.py:1077-1083: warn and use the batch size
            import warnings
            warnings.warn(f"Sample's batch dimension size {n} is different from the provided {num_samples} num_samples "
                          "argument. Defaulting to {n}.", UserWarning, stacklevel=2)
        self.model, self.posterior_samples = model, posterior_samples
        self.num_samples, self.return_sites, self.parallel = n, return_sites, parallel
        self.batch_ndims, self._batch_shape, self.exclude_deterministic = batch_ndims, tuple(shape), exclude_deterministic

    def __call__(self, rng_key, *args, **kwargs) -> Dict[str, np.ndarray]:
        rng_key = np.asarray(rng_key, np.uint32).reshape(2)
        n, shape = self.num_samples, self._batch_shape
        bound, e = _engine_for(self.model, args, kwargs)
        try:
            z = torch.from_numpy(_unconstrained(bound, self.posterior_samples, shape, e.D)).to(e.device)
            sample_keys = rng_key[None] if n <= 1 else b2random.split(rng_key, n)                  # util.py:916-918
            keys = b2random.split_each(sample_keys)[:, 1]                                          # seed handler, handlers.py:896
            n_obs = int(e.lib.b200nuts_obs_count(e.h))
            draws = _in_chunks(n, n_obs, lambda lo, hi: e.predict(z[lo:hi], keys[lo:hi]))
            con = e.constrain(z).cpu().numpy()
        finally:
            e.close()
        if bound.cfg.get("likelihood") == _capi.LIK_BERNOULLI_LOGIT and bound.cfg.get("family") == _capi.FAMILY_GLM:
            draws = draws.astype(np.int32)                                   # discrete.py:226-241: bernoulli draws are integers
        out = {self.model.obs_name: draws.reshape(shape + (n_obs,))}
        for s in bound.sites:
            if s.deterministic:
                out[s.name] = con[:, s.c_offset:s.c_offset + s.size].reshape(shape + tuple(s.shape))
        if self.return_sites is not None and self.return_sites != "":
            latent = {s.name: np.asarray(self.posterior_samples[s.name]) for s in bound.latent_sites}
            allsites = dict(latent, **out)
            return {k: allsites[k] for k in self.return_sites if k in allsites}
        if self.return_sites == "":
            out.update({s.name: np.asarray(self.posterior_samples[s.name]) for s in bound.latent_sites})
        return out
