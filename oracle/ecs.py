"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the Gibbs-within-HMC callers of the hot path,
numpyro/infer/hmc_gibbs.py ``HMCGibbs`` :38-192 and ``HMCECS`` :502-690 with numpyro/contrib/ecs_proxies.py
(``perturbed_method`` :23-50, ``_update_block`` :58-71, ``taylor_proxy`` :95-300) for the plain GLM
(examples/covtype.py:66-71 with ``subsample_size``), on top of the oracle's own NUTS/HMC kernel (oracle/chain.py).

Arithmetic that lives in the un-vendored dependency jax (``jax.random.randint``, used by ``_update_block`` and by the CPU
branch of ``_subsample_fn``, numpyro/primitives.py:457-469) is restated from jax's published algorithm (two 32-bit draws,
``span``-modular combination) -- "parity unpinned": the reference holds no vector for it.
The potential is evaluated in fp64 and rounded once, like every oracle potential; the Taylor terms jax obtains by automatic
differentiation are written out for a GLM (the row log-likelihood depends on the parameters through eta = x . z only).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
from scipy.special import gammaln

from . import chain as ch
from . import detmath as dm
from . import prng

F = np.float32
U32 = np.uint32


# ---------------------------------------------------------------- jax.random.randint (int32, partitionable threefry)
def randint(key, n: Optional[int], minval: int, maxval: int) -> np.ndarray:
    """``jax.random.randint(key, shape, minval, maxval)`` for int32 and 0 <= minval < maxval <= 2^31 - 1: two draws of 32 bits
    from ``split(key)``, combined as ((hi % span) * (2^32 % span) + lo % span) % span in wrapping uint32 arithmetic."""
    k1, k2 = prng.split(np.asarray(key, U32))
    hi = np.atleast_1d(prng.random_bits(k1, n)).astype(np.uint64)
    lo = np.atleast_1d(prng.random_bits(k2, n)).astype(np.uint64)
    span = np.uint64(max(int(maxval) - int(minval), 1))
    mask = np.uint64(0xFFFFFFFF)
    mult = (np.uint64(1) << np.uint64(16)) % span
    mult = ((mult * mult) & mask) % span
    off = ((((hi % span) * mult) & mask) + (lo % span)) & mask
    off = off % span
    out = (np.int64(minval) + off.astype(np.int64)).astype(np.int32)
    return out[0] if n is None else out


def subsample_indices(key, size: int, m: int) -> np.ndarray:
    """numpyro/primitives.py:457-469 (CPU branch of ``_subsample_fn``): partial Fisher-Yates from the back of arange(size)."""
    keys = prng.split(np.asarray(key, U32), m)
    val = np.arange(size, dtype=np.int32)
    for idx in range(m):
        i_p1 = size - idx
        i = i_p1 - 1
        j = int(randint(keys[idx], None, 0, i_p1))
        val[i], val[j] = val[j], val[i]
    return val[-m:].copy()


def update_block(key, num_blocks: int, idx: np.ndarray, size: int):
    """contrib/ecs_proxies.py:58-71.  Returns (next key, new subsample indices)."""
    m = idx.shape[0]
    key, subkey, block_key = prng.split(np.asarray(key, U32), 3)
    block_size = (m - 1) // num_blocks + 1
    pad = block_size - (m - 1) % block_size - 1
    chosen = int(randint(block_key, None, 0, num_blocks))
    new_idx = randint(subkey, block_size, 0, size)
    padded = np.concatenate([idx, np.zeros(pad, np.int32)])
    start = chosen * block_size
    padded[start:start + block_size] = new_idx
    return key, padded[:m].astype(np.int32)


# ---------------------------------------------------------------- Taylor proxy for a plain GLM
def _row_loglik(lik, eta, y):
    """(l, dl/deta, d2l/deta2) of one row's log-likelihood (discrete.py:263, :1388)."""
    with np.errstate(all="ignore"):
        if lik == "bernoulli":
            s = 1.0 / (1.0 + np.exp(-eta))
            return -(np.maximum(eta, 0) + np.log1p(np.exp(-np.abs(eta))) - eta * y), y - s, -s * (1 - s)
        r = np.exp(eta)
        return y * eta - r - gammaln(y + 1.0), y - r, -r


@dataclass
class TaylorProxy:
    """taylor_proxy (contrib/ecs_proxies.py:95-300) at ``ref`` (unconstrained = constrained for ``coefs``)."""
    ref: np.ndarray
    degree: int
    eta_ref: np.ndarray       # [N]
    L0: float                 # ref_sum_log_lik
    G: np.ndarray             # ref_sum_log_lik_grads [D]
    H: np.ndarray             # ref_sum_log_lik_hessians [D, D]

    @staticmethod
    def build(X, y, lik, ref, degree=2) -> "TaylorProxy":
        X64, y64, ref = np.asarray(X, np.float64), np.asarray(y, np.float64), np.asarray(ref, np.float64)
        e0 = X64 @ ref
        l, d1, d2 = _row_loglik(lik, e0, y64)
        return TaylorProxy(ref, degree, e0, float(l.sum()), X64.T @ d1, (X64 * d2[:, None]).T @ X64)


def ecs_potential64(X, y, lik, z, idx, proxy: Optional[TaylorProxy]):
    """Potential of the inner kernel for the subsample ``idx``: -(log prior + estimate_likelihood factor).  Without a proxy the
    subsampled plate scales the likelihood by N / m (primitives.py plate); with it, perturbed_method (ecs_proxies.py:23-50)."""
    z = np.asarray(z, np.float64)
    N, m = X.shape[0], idx.shape[0]
    Xs, ys = np.asarray(X[idx], np.float64), np.asarray(y[idx], np.float64)
    eta = Xs @ z
    l, d1, _ = _row_loglik(lik, eta, ys)
    prior = np.sum(-0.5 * z * z - 0.5 * np.log(2 * np.pi))
    if proxy is None:
        ll = N / m * l.sum()
        g_ll = N / m * (Xs.T @ d1)
    else:
        dz = z - proxy.ref
        e0 = proxy.eta_ref[idx]
        l0, d10, d20 = _row_loglik(lik, e0, ys)
        a = eta - e0
        prox = l0 + d10 * a
        c = d1 - d10
        all_ = proxy.L0 + proxy.G @ dz
        g_all = proxy.G.copy()
        if proxy.degree == 2:
            prox = prox + 0.5 * d20 * a * a
            c = c - d20 * a
            all_ = all_ + 0.5 * dz @ proxy.H @ dz
            g_all = g_all + proxy.H @ dz
        diff = l - prox
        mean, var = diff.mean(), diff.var()
        ll = all_ + N * mean - 0.5 * (N * N / m) * var
        w = N / m - (N * N / (m * m)) * (diff - mean)
        g_ll = g_all + Xs.T @ (w * c)
    return -(prior + ll), -(-z + g_ll)


# ---------------------------------------------------------------- HMCECS (hmc_gibbs.py:502-690)
@dataclass
class ECSState:
    u: np.ndarray             # subsample indices (the Gibbs site)
    hmc_state: ch.HMCState
    rng_key: np.ndarray
    accept_prob: np.float32


class HMCECS:
    """One chain of HMCECS over the oracle kernel.  ``potential_at(u)`` returns the callable z -> (U, grad) for a subsample
    (in the GPU tests: the engine's own potential hook, so that the bookkeeping can be compared bit for bit)."""

    def __init__(self, kernel_kwargs, potential_at, size: int, m: int, num_blocks: int = 1):
        self.kw, self.potential_at, self.size, self.m, self.num_blocks = dict(kernel_kwargs), potential_at, size, m, num_blocks

    def _kernel(self, u):
        k = ch.Kernel(self.potential_at(u), **self.kw)
        if hasattr(self, "_adapter"):
            k.adapter, k.num_warmup = self._adapter, self._num_warmup
        return k

    def init(self, rng_key, num_warmup: int, family, init_z=None, has_proxy=True) -> ECSState:
        """HMCECS.init :577-638 -> HMCGibbs.init :123-151 -> HMC.init (hmc.py:740-799)."""
        rng_key, key_u = prng.split(np.asarray(rng_key, U32))
        k_after_coefs = prng.split(key_u)[0]                     # seed handler: the latent site takes split(key_u)[1] ...
        k_plate = prng.split(k_after_coefs)[1]                   # ... the subsample plate the next one (handlers.py:887-897)
        u = subsample_indices(k_plate, self.size, self.m)
        if has_proxy:
            rng_key, _rng_state = prng.split(rng_key)            # :627-628 (gibbs_init ignores its key)
        rng_key, key_z = prng.split(rng_key)                     # HMCGibbs.init :133
        k_hmc, k_init = prng.split(key_z)                        # hmc.py:744-750
        pot = self.potential_at(u)
        if init_z is None:
            z, pe, g, ok = ch.init_to_uniform(k_init, family.init_sites, family.layout, pot)
            assert ok
        else:
            z = np.asarray(init_z, F)
            pe, g = pot(z)
        kern = ch.Kernel(pot, **self.kw)
        hs = kern.init(k_hmc, num_warmup, z, pe, g)
        self._adapter, self._num_warmup = kern.adapter, num_warmup
        return ECSState(u, hs, rng_key, F(0.0))

    def sample(self, s: ECSState) -> ECSState:
        """HMCECS.sample :640-682 (the update and the accept draw both consume ``rng_key``; ``rng_gibbs`` is unused there)."""
        rng_key, _rng_gibbs = prng.split(s.rng_key)
        _, u_new = update_block(rng_key, self.num_blocks, s.u, self.size)
        pe = s.hmc_state.potential_energy
        pe_new, g_new = self.potential_at(u_new)(s.hmc_state.z)
        with np.errstate(all="ignore"):
            acc = dm.exp(F(F(pe) - F(pe_new)))
            acc = F(1.0) if acc > F(1.0) else acc
        take = bool(prng.uniform(rng_key) < acc)
        u, g = (u_new, np.asarray(g_new, F)) if take else (s.u, s.hmc_state.z_grad)
        pe = F(pe_new) if take else pe
        hs = replace(s.hmc_state, z_grad=g, potential_energy=F(pe))
        hs = self._kernel(u).sample(hs)
        return ECSState(u, hs, rng_key, acc)


# ---------------------------------------------------------------- HMCGibbs (hmc_gibbs.py:38-192)
@dataclass
class GibbsState:
    gibbs: dict               # Gibbs sites (constrained values)
    hmc_state: ch.HMCState
    rng_key: np.ndarray


class HMCGibbs:
    """One chain of HMC-within-Gibbs over the oracle kernel.  ``potential_at(gibbs)`` -> (z_hmc -> (U, grad)) is the potential
    conditioned on the Gibbs sites, ``constrain_hmc(z_hmc, gibbs)`` the inner kernel's postprocess_fn, ``prior_draw(key_u)`` the
    Gibbs sites' values in the seeded prototype trace (:127-131)."""

    def __init__(self, kernel_kwargs, potential_at, gibbs_fn, constrain_hmc, prior_draw):
        self.kw, self.potential_at, self.gibbs_fn = dict(kernel_kwargs), potential_at, gibbs_fn
        self.constrain_hmc, self.prior_draw = constrain_hmc, prior_draw

    def init(self, rng_key, num_warmup: int, reduced_family, init_z=None) -> GibbsState:
        rng_key, key_u = prng.split(np.asarray(rng_key, U32))
        gibbs = self.prior_draw(key_u)
        rng_key, key_z = prng.split(rng_key)
        k_hmc, k_init = prng.split(key_z)
        pot = self.potential_at(gibbs)
        if init_z is None:
            z, pe, g, ok = ch.init_to_uniform(k_init, reduced_family.init_sites, reduced_family.layout, pot)
            assert ok
        else:
            z = np.asarray(init_z, F)
            pe, g = pot(z)
        kern = ch.Kernel(pot, **self.kw)
        hs = kern.init(k_hmc, num_warmup, z, pe, g)
        self._adapter, self._num_warmup = kern.adapter, num_warmup
        return GibbsState(gibbs, hs, rng_key)

    def sample(self, s: GibbsState) -> GibbsState:
        rng_key, rng_gibbs = prng.split(s.rng_key)
        z_hmc = self.constrain_hmc(s.hmc_state.z, s.gibbs)
        gibbs = self.gibbs_fn(rng_key=rng_gibbs, gibbs_sites=s.gibbs, hmc_sites=z_hmc)
        pot = self.potential_at(gibbs)
        pe, g = pot(s.hmc_state.z)
        hs = replace(s.hmc_state, z_grad=np.asarray(g, F), potential_energy=F(pe))
        k = ch.Kernel(pot, **self.kw)
        k.adapter, k.num_warmup = self._adapter, self._num_warmup
        return GibbsState(gibbs, k.sample(hs), rng_key)
