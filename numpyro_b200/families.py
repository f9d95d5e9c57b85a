"""Registered model families: the host-side description the engine needs in place of a traced model.

In numpyro the model is an arbitrary Python callable that ``initialize_model``
(numpyro/infer/util.py:663-835) traces to find the latent sites, their supports and the potential.
The arithmetic between sites is opaque to a trace, so the engine instead takes a *declared* family
(SURVEY.md 7.2 item 8): each class below names the numpyro model it stands for, lists its latent
sites in model-trace order with their constraints, and maps the ``mcmc.run(key, *args, **kwargs)``
arguments of that model to the engine's data pointers.  Anything else raises -- there is no
fallback path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _capi


@dataclass
class Site:
    name: str
    shape: Tuple[int, ...]
    positive: bool = False           # support = positive -> ExpTransform (transforms.py:2115-2118)
    deterministic: bool = False      # numpyro.deterministic site produced by postprocess_fn
    z_offset: int = -1               # offset in the flat unconstrained vector (latent sites)
    c_offset: int = -1               # offset in one row of b200nuts_constrain output

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1


@dataclass
class BoundModel:
    """A family bound to its data: everything ``Engine(**cfg, X=, y=, aux=)`` needs."""
    cfg: Dict
    X: Optional[np.ndarray]
    y: Optional[np.ndarray]
    aux: Optional[np.ndarray]
    sites: List[Site]                # latent sites in flat (sorted-name) order, then deterministic sites
    trace_order: Optional[List[str]] = None     # latent site names in model-trace order (None: the flat order)

    @property
    def latent_sites(self) -> List[Site]:
        return [s for s in self.sites if not s.deterministic]


def _layout(latent: List[Site], deterministic: List[Site]) -> List[Site]:
    off = 0
    latent = sorted(latent, key=lambda s: s.name)          # hmc.py:765-768: sorted site names
    for s in latent:
        s.z_offset = s.c_offset = off
        off += s.size
    for s in deterministic:
        s.c_offset = off
        off += s.size
    return latent + deterministic


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Model:
    """Base class of declared families; ``bind`` receives the model's call arguments.  ``obs_name`` is the observed site
    (what ``log_likelihood`` / ``Predictive`` return); ``bind(..., _predict=True)`` accepts ``y=None`` (the reference's
    convention for predictive calls, infer/util.py:985-998) and binds a placeholder response."""
    name = "model"
    obs_name = "obs"

    def bind(self, *args, **kwargs) -> BoundModel:
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        raise TypeError(f"{type(self).__name__} is a declared model family for numpyro_b200's NUTS/HMC engine; "
                        "it is not executable as a numpyro program")


class DiagGaussian(Model):
    """``x ~ Normal(mu, sigma)`` elementwise; analytic target used by the tests."""
    name = "diag_gaussian"

    def __init__(self, mu, sigma):
        self.mu, self.sigma = _f32(mu).ravel(), _f32(sigma).ravel()

    def bind(self) -> BoundModel:
        d = self.mu.shape[0]
        return BoundModel(dict(family=_capi.FAMILY_DIAG_GAUSSIAN, n_rows=d), None, None,
                          np.concatenate([self.mu, self.sigma]), _layout([Site("x", (d,))], []))


class EightSchoolsNonCentered(Model):
    """README.md:100-110::

        def eight_schools_noncentered(J, sigma, y=None):
            mu = sample('mu', Normal(0, 5)); tau = sample('tau', HalfCauchy(5))
            with plate('J', J), reparam(config={'theta': TransformReparam()}):
                theta = sample('theta', TransformedDistribution(Normal(0, 1), AffineTransform(mu, tau)))
            sample('obs', Normal(theta, sigma), obs=y)
    """
    name = "eight_schools_noncentered"

    def __init__(self, mu_scale: float = 5.0, tau_scale: float = 5.0):
        self.mu_scale, self.tau_scale = float(mu_scale), float(tau_scale)

    def bind(self, J, sigma, y=None, _predict=False) -> BoundModel:
        if y is None:
            if not _predict:
                raise ValueError("the engine samples posteriors: `y` must be observed")
            y = np.zeros(int(J), np.float32)
        sigma, y = _f32(sigma).ravel(), _f32(y).ravel()
        if sigma.shape[0] != J or y.shape[0] != J:
            raise ValueError("sigma and y must have length J")
        sites = _layout([Site("mu", ()), Site("tau", (), positive=True), Site("theta_base", (J,))],
                        [Site("theta", (J,), deterministic=True)])
        cfg = dict(family=_capi.FAMILY_EIGHT_SCHOOLS, n_rows=int(J), tau_scale=self.tau_scale, mu_scale=self.mu_scale)
        return BoundModel(cfg, None, y, sigma, sites)


class _GLM(Model):
    likelihood = _capi.LIK_BERNOULLI_LOGIT
    coef_name = "coefs"

    def _cfg(self, X) -> Dict:
        return dict(family=_capi.FAMILY_GLM, likelihood=self.likelihood)

    def _sites(self, D) -> Tuple[List[Site], List[Site]]:
        return [Site(self.coef_name, (D,))], []

    def bind(self, X, y=None, _predict=False) -> BoundModel:
        if y is None:
            if not _predict:
                raise ValueError("the engine samples posteriors: `y` must be observed")
            y = np.zeros(np.shape(X)[0], np.float32)
        X, y = _f32(X), _f32(y).ravel()
        if X.ndim != 2 or y.shape[0] != X.shape[0]:
            raise ValueError("X must be [N, D] and y [N]")
        lat, det = self._sites(X.shape[1])
        return BoundModel(self._cfg(X), X, y, None, _layout(lat, det), trace_order=[s.name for s in lat])


class LogisticRegression(_GLM):
    """examples/covtype.py:66-71: ``coefs ~ Normal(0, 1)[D]``, ``obs ~ Bernoulli(logits=data @ coefs)``."""
    name = "logistic_regression"


class PoissonRegression(_GLM):
    """``coefs ~ Normal(0, 1)[D]``, ``obs ~ Poisson(exp(X @ coefs))`` (discrete.py:1361-1388)."""
    name = "poisson_regression"
    likelihood = _capi.LIK_POISSON_LOG


class HorseshoeRegression(_GLM):
    """examples/horseshoe_regression.py:37-78 (``model_normal_likelihood`` / ``model_bernoulli_likelihood``)."""
    name = "horseshoe_regression"
    coef_name = "unscaled_betas"
    obs_name = "Y"

    def __init__(self, likelihood: str = "normal"):
        if likelihood not in ("normal", "bernoulli"):
            raise ValueError("likelihood must be 'normal' or 'bernoulli'")
        self.likelihood = _capi.LIK_NORMAL if likelihood == "normal" else _capi.LIK_BERNOULLI_LOGIT

    def _cfg(self, X):
        return dict(family=_capi.FAMILY_GLM, likelihood=self.likelihood, local_scales=1,
                    global_scale=_capi.SCALE_HALFCAUCHY, tau_scale=1.0)

    def _sites(self, D):
        lat = [Site("lambdas", (D,), positive=True), Site("tau", (1,), positive=True), Site("unscaled_betas", (D,))]
        if self.likelihood == _capi.LIK_NORMAL:
            lat.append(Site("prec_obs", (), positive=True))
        return lat, [Site("betas", (D,), deterministic=True)]


class HierarchicalGLM(_GLM):
    """BASELINE config 3: GLM whose group columns [g0, g1) carry non-centred random effects::

        tau ~ HalfCauchy(tau_scale) | Exponential(1/tau_scale);  coefs ~ Normal(0, 1)[D]
        betas = coefs * where(g0 <= j < g1, tau, 1);  obs ~ Bernoulli(logits=X @ betas) | Poisson(exp(.))
    """
    name = "hierarchical_glm"

    def __init__(self, group_cols: Tuple[int, int], likelihood: str = "bernoulli", tau_prior: str = "halfcauchy",
                 tau_scale: float = 1.0):
        self.group_cols = (int(group_cols[0]), int(group_cols[1]))
        self.likelihood = {"bernoulli": _capi.LIK_BERNOULLI_LOGIT, "poisson": _capi.LIK_POISSON_LOG}[likelihood]
        self.gscale = {"halfcauchy": _capi.SCALE_HALFCAUCHY, "exponential": _capi.SCALE_EXPONENTIAL}[tau_prior]
        self.tau_scale = float(tau_scale)

    def _cfg(self, X):
        return dict(family=_capi.FAMILY_GLM, likelihood=self.likelihood, global_scale=self.gscale,
                    group_col_begin=self.group_cols[0], group_col_end=self.group_cols[1], tau_scale=self.tau_scale)

    def _sites(self, D):
        return [Site("tau", (1,), positive=True), Site("coefs", (D,))], [Site("betas", (D,), deterministic=True)]
