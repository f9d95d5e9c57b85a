"""Oracle: potentials of the registered model families (TEST INFRASTRUCTURE ONLY).

Each family restates ``potential_energy`` (numpyro/infer/util.py:333-358): minus the sum of the
site log-probs plus the log|det J| factors that ``_unconstrain_reparam`` (:300-330) injects for
constrained sites.  The log-prob formulas followed are numpyro/distributions/continuous.py
Normal :2975-2989, Cauchy :390-406, HalfCauchy :1148-1150, Exponential :715-726, Gamma :813-831;
discrete.py BernoulliLogits :263 (-> distributions/util.py:317-320), Poisson :1361-1388;
transforms.py ExpTransform :635-646; distribution.py TransformedDistribution :1277-1302 with
infer/reparam.py TransformReparam :166-191 for the non-centred eight-schools model.

U and grad(U) are evaluated in float64 with hand-derived gradients and rounded ONCE to float32
(SURVEY.md 7.3); ``tests/test_oracle_families.py`` checks every gradient by central differences.
The flat layout is numpyro's: site names sorted, each raveled row-major (hmc.py:765-768);
``init_sites`` lists the latent sites in MODEL-TRACE order, which is the order
``find_valid_initial_params`` consumes PRNG keys in (infer/util.py:454-463).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
from scipy.special import gammaln

F = np.float32
LOG_SQRT_2PI = 0.5 * np.log(2 * np.pi)


def _softplus_neg_abs(eta):
    return np.log1p(np.exp(-np.abs(eta)))


class Family:
    name = "family"
    init_sites: List[Tuple[str, int]]            # (site, size) in model-trace order

    @property
    def layout(self) -> List[Tuple[str, int, int]]:
        """(site, offset, size) in flat (sorted-name) order."""
        out, off = [], 0
        for n, s in sorted(self.init_sites):
            out.append((n, off, s))
            off += s
        return out

    @property
    def dim(self) -> int:
        return sum(s for _, s in self.init_sites)

    def potential64(self, z64):
        raise NotImplementedError

    def potential_and_grad(self, z):
        with np.errstate(all="ignore"):
            u, g = self.potential64(np.asarray(z, np.float64))
        return F(u), np.asarray(g, F)

    def constrain(self, z) -> Dict[str, np.ndarray]:
        """postprocess_fn: constrained latent sites + deterministic sites (mcmc.py:193-214)."""
        raise NotImplementedError


class DiagGaussian(Family):
    """U = 0.5 * sum((z - mu)^2 / sigma^2): analytic test target (test_hmc_util.py style)."""
    name = "diag_gaussian"

    def __init__(self, mu, sigma):
        self.mu = np.asarray(mu, np.float64)
        self.sigma = np.asarray(sigma, np.float64)
        self.init_sites = [("x", self.mu.shape[0])]

    def potential64(self, z):
        d = (z - self.mu) / self.sigma
        return 0.5 * np.sum(d * d), d / self.sigma

    def constrain(self, z):
        return {"x": np.asarray(z, F)}


class EightSchools(Family):
    """Non-centred eight schools (README.md:100-110); latent mu, tau, theta_base[J]."""
    name = "eight_schools"

    def __init__(self, sigma, y, mu_scale=5.0, tau_scale=5.0):
        self.sigma = np.asarray(sigma, np.float64)
        self.y = np.asarray(y, np.float64)
        self.J = self.y.shape[0]
        self.mu_scale, self.tau_scale = float(mu_scale), float(tau_scale)
        self.init_sites = [("mu", 1), ("tau", 1), ("theta_base", self.J)]

    def potential64(self, z):
        mu, zt, tb = z[0], z[1], z[2:]
        tau = np.exp(zt)
        theta = mu + tau * tb
        resid = (self.y - theta) / self.sigma
        lp = -0.5 * (mu / self.mu_scale) ** 2 - np.log(self.mu_scale) - LOG_SQRT_2PI
        q = (tau / self.tau_scale) ** 2
        lp += np.log(2.0) - np.log(np.pi) - np.log(self.tau_scale) - np.log1p(q) + zt
        lp += np.sum(-0.5 * tb * tb - LOG_SQRT_2PI)
        lp += np.sum(-0.5 * resid * resid - np.log(self.sigma) - LOG_SQRT_2PI)
        d_theta = resid / self.sigma                   # d lp / d theta_j
        g = np.empty_like(z)
        g[0] = -(-mu / self.mu_scale ** 2 + np.sum(d_theta))
        g[1] = -(-2.0 * q / (1.0 + q) + 1.0 + np.sum(d_theta * tb) * tau)
        g[2:] = -(-tb + d_theta * tau)
        return -lp, g

    def constrain(self, z):
        z = np.asarray(z, F)
        with np.errstate(all="ignore"):
            tau = np.exp(z[..., 1])
            theta = z[..., 0:1] + tau[..., None] * z[..., 2:]
        return {"mu": z[..., 0], "tau": tau.astype(F), "theta_base": z[..., 2:],
                "theta": theta.astype(F)}


class GLM(Family):
    """Generalised linear model with (optionally shrunk) coefficients -- one class for:

    * plain GLM, covtype model (examples/covtype.py:66-71): coefs ~ N(0,1), obs ~ Bernoulli(logits)
      or Poisson(exp(eta));
    * horseshoe regression (examples/horseshoe_regression.py:37-78): lambdas ~ HalfCauchy(1)[D],
      tau ~ HalfCauchy(1)[1], unscaled_betas ~ N(0,1)[D], betas = tau*lambdas*unscaled_betas,
      Bernoulli-logit likelihood, or Normal likelihood with prec_obs ~ Gamma(3,1);
    * hierarchical GLM (BASELINE config 3): a global scale tau ~ HalfCauchy/Exponential multiplies
      the coefficients of the group columns [g0, g1) (non-centred random effects).
    """

    def __init__(self, X, y, likelihood="bernoulli", local_scales=False, global_scale=None,
                 group_cols=None, tau_scale=1.0, coef_name="coefs"):
        self.X = np.asarray(X, np.float64)
        self.y = np.asarray(y, np.float64)
        self.N, self.D = self.X.shape
        assert likelihood in ("bernoulli", "poisson", "normal")
        assert global_scale in (None, "halfcauchy", "exponential")
        self.likelihood, self.local, self.gscale = likelihood, local_scales, global_scale
        self.g0, self.g1 = group_cols if group_cols is not None else (0, self.D)
        self.tau_scale = float(tau_scale)
        self.coef_name = coef_name
        sites = []
        if self.local:
            sites.append(("lambdas", self.D))
        if self.gscale:
            sites.append(("tau", 1))
        sites.append((coef_name, self.D))
        if likelihood == "normal":
            sites.append(("prec_obs", 1))
        self.init_sites = sites
        self.name = "glm_" + likelihood
        self._lgam = gammaln(self.y + 1.0) if likelihood == "poisson" else None

    def _split(self, z):
        out = {}
        for n, off, s in self.layout:
            out[n] = z[..., off:off + s]
        return out

    def coef_scale(self, parts):
        s = np.ones(self.D)
        if self.local:
            s = s * np.exp(parts["lambdas"])
        if self.gscale:
            t = np.ones(self.D)
            t[self.g0:self.g1] = np.exp(parts["tau"][0])
            s = s * t
        return s

    def potential64(self, z):
        p = self._split(z)
        u = p[self.coef_name]
        s = self.coef_scale(p)
        beta = s * u
        eta = self.X @ beta
        if self.likelihood == "bernoulli":
            ll = -np.sum(np.maximum(eta, 0) + _softplus_neg_abs(eta) - eta * self.y)
            dl = self.y - 1.0 / (1.0 + np.exp(-eta))            # d ll / d eta
        elif self.likelihood == "poisson":
            ll = np.sum(self.y * eta - self._lgam - np.exp(eta))
            dl = self.y - np.exp(eta)
        else:
            zp = p["prec_obs"][0]
            prec = np.exp(zp)
            res = self.y - eta
            ss = np.sum(res * res)
            ll = -0.5 * prec * ss - self.N * (-0.5 * zp) - self.N * LOG_SQRT_2PI
            dl = prec * res
        g_beta = self.X.T @ dl                                   # d ll / d beta
        lp = ll + np.sum(-0.5 * u * u - LOG_SQRT_2PI)
        g = {self.coef_name: -u + s * g_beta}
        if self.local:
            zl = p["lambdas"]
            lam2 = np.exp(2.0 * zl)
            lp += np.sum(np.log(2.0) - np.log(np.pi) - np.log1p(lam2) + zl)
            g["lambdas"] = beta * g_beta - 2.0 * lam2 / (1.0 + lam2) + 1.0
        if self.gscale:
            zt = p["tau"][0]
            tau = np.exp(zt)
            if self.gscale == "halfcauchy":
                q = (tau / self.tau_scale) ** 2
                lp += np.log(2.0) - np.log(np.pi) - np.log(self.tau_scale) - np.log1p(q) + zt
                dprior = -2.0 * q / (1.0 + q) + 1.0
            else:                                                # Exponential(rate = 1/tau_scale)
                rate = 1.0 / self.tau_scale
                lp += np.log(rate) - rate * tau + zt
                dprior = -rate * tau + 1.0
            g["tau"] = np.array([np.sum((beta * g_beta)[self.g0:self.g1]) + dprior])
        if self.likelihood == "normal":
            # Gamma(3, 1) on prec (continuous.py:825-831) + exp-transform Jacobian
            lp += 2.0 * zp - prec - gammaln(3.0) + zp
            g["prec_obs"] = np.array([-0.5 * prec * ss + 0.5 * self.N + 3.0 - prec])
        gz = np.empty_like(z)
        for n, off, sz in self.layout:
            gz[off:off + sz] = -g[n]
        return -lp, gz

    def potential_and_grad_f32(self, z):
        """Plain-GLM potential in float32 with BLAS matvecs: the arithmetic a CPU run of the
        reference performs (fp32 ``jnp.dot(data, coefs)`` forward + transpose product backward).
        Used only as the timed CPU baseline of bench.py; parity tests use the fp64 path."""
        assert not self.local and not self.gscale and self.likelihood in ("bernoulli", "poisson")
        if not hasattr(self, "_X32"):
            self._X32 = np.ascontiguousarray(self.X, F)
            self._y32 = np.ascontiguousarray(self.y, F)
            self._lg32 = F(0.0) if self._lgam is None else F(np.sum(self._lgam))
        u = np.asarray(z, F)
        eta = self._X32 @ u
        with np.errstate(all="ignore"):
            if self.likelihood == "bernoulli":
                e = np.exp(-np.abs(eta))
                nll = np.sum(np.maximum(eta, 0) + np.log1p(e) - eta * self._y32, dtype=np.float64)
                dl = F(1.0) / (F(1.0) + np.exp(-eta)) - self._y32
            else:
                r = np.exp(eta)
                nll = np.sum(r - self._y32 * eta, dtype=np.float64) + self._lg32
                dl = r - self._y32
        g = u + self._X32.T @ dl
        U = nll + 0.5 * float(u @ u) + self.D * LOG_SQRT_2PI
        return F(U), g.astype(F)

    def constrain(self, z):
        z = np.asarray(z, F)
        p = self._split(z)
        out = {}
        with np.errstate(all="ignore"):
            if self.local:
                out["lambdas"] = np.exp(p["lambdas"]).astype(F)
            if self.gscale:
                out["tau"] = np.exp(p["tau"]).astype(F)
            out[self.coef_name] = p[self.coef_name]
            if self.likelihood == "normal":
                out["prec_obs"] = np.exp(p["prec_obs"]).astype(F)
            if self.local or self.gscale:
                s = np.ones(z.shape[:-1] + (self.D,), F)
                if self.local:
                    s = s * out["lambdas"]
                if self.gscale:
                    t = np.ones_like(s)
                    t[..., self.g0:self.g1] = out["tau"]
                    s = s * t
                out["betas"] = (s * p[self.coef_name]).astype(F)
        return out


def horseshoe(X, y, likelihood="bernoulli") -> GLM:
    """examples/horseshoe_regression.py:37-78.  Model-trace order of the latent sites:
    lambdas, tau, unscaled_betas, (prec_obs)."""
    fam = GLM(X, y, likelihood=likelihood, local_scales=True, global_scale="halfcauchy",
              coef_name="unscaled_betas")
    fam.name = "horseshoe_" + likelihood
    return fam


def logistic_regression(X, y) -> GLM:
    """examples/covtype.py:66-71."""
    return GLM(X, y, likelihood="bernoulli", coef_name="coefs")
