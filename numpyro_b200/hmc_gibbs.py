"""Gibbs-within-HMC callers of the engine: ``HMCGibbs`` and ``HMCECS``.

Mirrors numpyro/infer/hmc_gibbs.py -- ``HMCGibbs`` :38-192 (user Gibbs sampler for some sites, NUTS/HMC for the rest) and
``HMCECS`` :502-690 (HMC with energy-conserving subsampling: the Gibbs site is the index set of a subsampled plate, the
inner potential a bias-corrected estimate of the log-likelihood with a Taylor-proxy control variate,
numpyro/contrib/ecs_proxies.py) -- as ``MCMCKernel``s over the engine's ``init`` / ``transition`` entry points:

* the inner NUTS/HMC transition, its adaptation and the subsampled potential run in the CUDA engine (``csrc/tick.cuh``,
  ``potential_ecs_inwarp`` / the conditioned potential in ``csrc/families.cuh``);
* the outer step -- key splits, the block update of the subsample indices, the Metropolis test of the new subsample, the
  user's ``gibbs_fn`` -- is the reference's host-side control flow, restated here; its PRNG words come from the engine's
  Threefry (``b200nuts_prng_*``), its one transcendental (``exp`` of the energy difference) from ``b200nuts_detmath``.

``HMCECS`` is implemented for the plain GLM families (``LogisticRegression`` / ``PoissonRegression`` called with
``subsample_size``, examples/covtype.py:66-71 and :154-165).
"""
from __future__ import annotations

from collections import namedtuple
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import _capi, engine as _engine, families, random as b2random
from .infer import HMC, _KernelBase, _flatten_init, _state_from_engine

HMCGibbsState = namedtuple("HMCGibbsState", "z, hmc_state, rng_key")
HMCECSState = namedtuple("HMCECSState", "z, hmc_state, rng_key, gibbs_state, accept_prob")

U32 = np.uint32


# ------------------------------------------------------------------------------------------ integer draws (jax.random.randint)
def _randint(keys, n: Optional[int], minval: int, maxval) -> np.ndarray:
    """``vmap(jax.random.randint)`` over ``keys`` [K, 2] for int32: two 32-bit draws from ``split(key)``, combined as
    ((hi % span) * (2^32 % span) + lo % span) % span in wrapping uint32 arithmetic.  ``maxval``: scalar or [K]."""
    keys = np.asarray(keys, U32).reshape(-1, 2)
    k12 = b2random.split_each(keys)                                          # [K, 2, 2]
    m = 1 if n is None else int(n)
    hi = np.stack([_engine.prng_bits(k, m) for k in k12[:, 0]]).astype(np.uint64)
    lo = np.stack([_engine.prng_bits(k, m) for k in k12[:, 1]]).astype(np.uint64)
    span = np.maximum(np.asarray(maxval, np.int64) - int(minval), 1).astype(np.uint64).reshape(-1, 1)
    mask = np.uint64(0xFFFFFFFF)
    mult = (np.uint64(1) << np.uint64(16)) % span
    mult = ((mult * mult) & mask) % span
    off = ((((hi % span) * mult) & mask) + (lo % span)) & mask
    out = (int(minval) + (off % span).astype(np.int64)).astype(np.int32)
    return out[:, 0] if n is None else out


def _subsample_indices(key, size: int, m: int) -> np.ndarray:
    """numpyro/primitives.py:457-469 (the CPU branch of ``_subsample_fn``): partial Fisher-Yates from the back of arange(size)."""
    keys = b2random.split(key, m)
    js = _randint(keys, None, 0, size - np.arange(m))
    val = np.arange(size, dtype=np.int32)
    for idx in range(m):
        i = size - idx - 1
        j = int(js[idx])
        val[i], val[j] = val[j], val[i]
    return val[-m:].copy()


def _update_block(keys, num_blocks: int, idx: np.ndarray, size: int) -> np.ndarray:
    """contrib/ecs_proxies.py:58-71 for every chain: ``idx`` [C, m] -> new indices [C, m]."""
    C, m = idx.shape
    k3 = b2random.split_each(keys, 3)                                        # rng_key, subkey, block_key
    block_size = (m - 1) // num_blocks + 1
    pad = block_size - (m - 1) % block_size - 1
    chosen = _randint(k3[:, 2], None, 0, num_blocks)
    new_idx = _randint(k3[:, 1], block_size, 0, size)
    padded = np.concatenate([idx, np.zeros((C, pad), np.int32)], axis=1)
    for c in range(C):
        start = int(chosen[c]) * block_size
        padded[c, start:start + block_size] = new_idx[c]
    return np.ascontiguousarray(padded[:, :m])


# ------------------------------------------------------------------------------------------ Taylor proxy
class _TaylorProxy:
    def __init__(self, reference_params: Dict, degree: int):
        if degree not in (1, 2):
            raise ValueError("Taylor proxy only defined for first and second degree.")
        self.reference_params, self.degree = reference_params, degree


def taylor_proxy(reference_params, degree=2):
    """``HMCECS.taylor_proxy`` (hmc_gibbs.py:684-690, contrib/ecs_proxies.py:95-107): control variate from a Taylor expansion of
    every row's log-likelihood around ``reference_params`` (MLE / MAP)."""
    return _TaylorProxy(dict(reference_params), degree)


# ------------------------------------------------------------------------------------------ kernels
class HMCGibbs(_KernelBase):
    """HMC-within-Gibbs (hmc_gibbs.py:38-192): ``gibbs_fn(rng_key, gibbs_sites, hmc_sites)`` resamples ``gibbs_sites`` from
    their conditional, NUTS/HMC (the engine) updates the remaining sites given them."""
    sample_field = "z"

    def __init__(self, inner_kernel, gibbs_fn, gibbs_sites):
        if not isinstance(inner_kernel, HMC):
            raise ValueError("inner_kernel must be an HMC or NUTS sampler.")
        if not callable(gibbs_fn):
            raise ValueError("gibbs_fn must be a callable")
        self.inner_kernel = inner_kernel
        self._gibbs_fn, self._gibbs_sites = gibbs_fn, list(gibbs_sites or [])
        self._engine: Optional[_engine.Engine] = None
        self._bound = None
        self._num_warmup = 0
        self._single = True

    @property
    def model(self):
        return self.inner_kernel.model

    @property
    def default_fields(self):
        return ("z",)

    @property
    def is_ensemble_kernel(self):
        return False

    def get_diagnostics_str(self, state):
        return self.inner_kernel.get_diagnostics_str(state.hmc_state)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    # ---- shared plumbing: the inner kernel's engine, conditioned on the Gibbs sites
    def _make_engine(self, keys, num_warmup, bound, extra_cfg):
        inner = self.inner_kernel
        cfg = dict(inner._cfg)
        cfg.update(bound.cfg)
        cfg.update(extra_cfg)
        cfg["num_chains"] = keys.shape[0]
        if self._engine is not None:
            self._engine.close()
        e = _engine.Engine(device=torch.device("cuda", torch.cuda.current_device()), X=bound.X, y=bound.y, aux=bound.aux, **cfg)
        if inner._inverse_mass_matrix is not None:
            e.set_inverse_mass_matrix(inner._inverse_mass_matrix)
        self._engine, self._bound, self._num_warmup = e, bound, int(num_warmup)
        return e

    def _hmc_state(self):
        e = self._engine
        st, vec = e.state()
        return _state_from_engine([st], [vec], self._hmc_bound, self._single, self.inner_kernel._trajectory_length,
                                  dense=[e.dense_state()] if self.inner_kernel._dense_mass else None), st, vec

    def _inner_transition(self, st, vec, pe, z_grad):
        """``hmc_state._replace(z_grad=..., potential_energy=...)`` then ``inner_kernel.sample`` (hmc_gibbs.py:176-182)."""
        e = self._engine
        vec = dict(vec)
        vec["z_grad"] = np.ascontiguousarray(z_grad, np.float32)
        for k in range(e.C):
            st[k].potential_energy = float(pe[k])
        e.set_state(st, vec, self._num_warmup)
        if self.inner_kernel._dense_mass:
            e.set_dense_state(self._dense["inverse_mass_matrix"], self._dense["wf_m2"])
        e.transition(1)

    # ---- MCMCKernel
    def _layout(self, bound):
        """Full-model coordinates of the Gibbs sites, and the inner kernel's (reduced) site table."""
        names = {s.name for s in bound.latent_sites}
        missing = [g for g in self._gibbs_sites if g not in names]
        if missing:
            raise ValueError(f"gibbs_sites {missing} are not latent sites of the model ({sorted(names)})")
        Dfull = sum(s.size for s in bound.latent_sites)
        fixed = np.zeros(Dfull, np.int32)
        hmc_sites, off = [], 0
        for s in bound.latent_sites:                                          # flat (sorted-name) order
            if s.name in self._gibbs_sites:
                fixed[s.z_offset:s.z_offset + s.size] = 1
            else:
                hmc_sites.append(families.Site(s.name, s.shape, s.positive, False, off, off))
                off += s.size
        return fixed, families.BoundModel(bound.cfg, bound.X, bound.y, bound.aux, hmc_sites)

    def _prior_draw(self, site, key, bound):
        """The prototype trace's draw of a Gibbs site (hmc_gibbs.py:127-131: seed + init_to_sample), for the priors restated
        here: Normal(0, scale).  Other priors: pass the initial value through ``init_params``."""
        if bound.cfg.get("family") == _capi.FAMILY_EIGHT_SCHOOLS and site.name == "mu":
            return np.float32(bound.cfg.get("mu_scale", 5.0)) * _engine.prng_normal(key, 1)[0]          # continuous.py:2961-2967
        raise NotImplementedError(f"initial value of Gibbs site {site.name!r}: pass it in init_params "
                                  "(only Normal priors are drawn from the prototype trace here)")

    def _push_gibbs(self, gibbs_values):
        """gibbs_values: {site: [C, ...] constrained values} -> the conditioned handle."""
        e, bound = self._engine, self._bound
        vals = np.zeros((e.C, e.Dfull + 1), np.float32)
        for s in bound.latent_sites:
            if s.name in self._gibbs_sites:
                v = np.asarray(gibbs_values[s.name], np.float32).reshape(e.C, s.size)
                if s.positive:
                    v = np.log(v)
                    vals[:, -1] += v.sum(axis=1)                              # no Jacobian term for a conditioned site
                vals[:, s.z_offset:s.z_offset + s.size] = v
        e.cond_set_values(vals)
        self._cond_vals = vals

    def _full_z(self, z_hmc_flat):
        full = self._cond_vals[:, :-1].copy()
        full[:, self._fixed == 0] = z_hmc_flat
        return full

    def _constrained(self, z_hmc_flat):
        """inner_kernel.postprocess_fn with the Gibbs sites substituted (hmc_gibbs.py:112-121, :166-168)."""
        e, bound = self._engine, self._bound
        con = e.constrain(torch.from_numpy(self._full_z(z_hmc_flat)).to(e.device)).cpu().numpy()
        return {s.name: con[:, s.c_offset:s.c_offset + s.size].reshape((e.C,) + tuple(s.shape)) for s in bound.sites
                if s.name not in self._gibbs_sites}

    def init(self, rng_key, num_warmup, init_params=None, model_args=(), model_kwargs=None):
        """hmc_gibbs.py:123-151, for one key or a batch of keys [C, 2]."""
        keys = np.asarray(rng_key, U32)
        self._single = keys.ndim == 1
        keys = keys.reshape(-1, 2)
        C = keys.shape[0]
        bound = self.inner_kernel.model.bind(*model_args, **(model_kwargs or {}))
        self._fixed, self._hmc_bound = self._layout(bound)
        init_params = dict(init_params or {})
        k2 = b2random.split_each(keys)                                        # rng_key, key_u (prototype trace)
        rng, key_u = k2[:, 0], k2[:, 1]
        gibbs = {}
        # seed handler (handlers.py:887-897): every latent site of the trace takes split(key)[1] in trace order
        order = {}
        kk = key_u
        for name in (bound.trace_order or [s.name for s in bound.latent_sites]):
            sp = b2random.split_each(kk)
            kk, order[name] = sp[:, 0], sp[:, 1]
        for s in bound.latent_sites:
            if s.name in self._gibbs_sites:
                if s.name in init_params:
                    gibbs[s.name] = np.broadcast_to(np.asarray(init_params.pop(s.name), np.float32), (C,) + tuple(s.shape)).copy()
                else:
                    gibbs[s.name] = np.stack([self._prior_draw(s, order[s.name][c], bound) for c in range(C)]).reshape((C,) + tuple(s.shape))
        k2 = b2random.split_each(rng)                                         # rng_key, key_z
        rng, key_z = k2[:, 0], k2[:, 1]
        e = self._make_engine(keys, num_warmup, bound, dict(cond_fixed=self._fixed, regime=_capi.REGIME_WARP))
        self._push_gibbs(gibbs)
        z0 = None
        if init_params:
            z0 = _flatten_init(init_params, self._hmc_bound, e.D)
        elif self.inner_kernel._init_strategy.kind == "feasible":
            z0 = np.zeros((C, e.D), np.float32)
        e.init(key_z, int(num_warmup), z0)
        e.run(0, 0, fields=())
        hs, _, _ = self._hmc_state()
        self._gibbs = gibbs
        return self._pack(hs, rng)

    def _pack(self, hs, rng):
        sq = (lambda a: a[0]) if self._single else (lambda a: a)
        z = {**{k: sq(v) for k, v in self._gibbs.items()}, **hs.z}
        return HMCGibbsState(z, hs, sq(np.asarray(rng, U32)))

    def sample(self, state, model_args=(), model_kwargs=None):
        """hmc_gibbs.py:153-186, all chains of the handle in lock-step (``gibbs_fn`` is called once per chain)."""
        e = self._engine
        if e is None:
            raise RuntimeError("sample() called before init()")
        C = e.C
        k2 = b2random.split_each(np.asarray(state.rng_key, U32).reshape(C, 2))
        rng, rng_gibbs = k2[:, 0], k2[:, 1]
        hs, st, vec = self._hmc_state()
        if self.inner_kernel._dense_mass:
            self._dense = e.dense_state()
        z_hmc = self._constrained(vec["z"])
        new = {k: [] for k in self._gibbs}
        for c in range(C):
            out = self._gibbs_fn(rng_key=rng_gibbs[c], gibbs_sites={k: v[c] for k, v in self._gibbs.items()},
                                 hmc_sites={k: v[c] for k, v in z_hmc.items()})
            for k in new:
                new[k].append(np.asarray(out[k], np.float32))
        self._gibbs = {k: np.stack(v).reshape(self._gibbs[k].shape) for k, v in new.items()}
        self._push_gibbs(self._gibbs)
        pe, grad = e.potential_and_grad(vec["z"])                             # value_and_grad(potential_fn(z_gibbs))(hmc_state.z)
        self._inner_transition(st, vec, pe.cpu().numpy(), grad.cpu().numpy())
        hs, _, _ = self._hmc_state()
        return self._pack(hs, rng)

    def postprocess_fn(self, args=(), kwargs=None):
        """hmc_gibbs.py:112-121: Gibbs sites as they are, HMC sites constrained (+ the deterministic sites)."""
        def fn(z):
            bound = self._bound
            first = np.asarray(z[self._hmc_bound.latent_sites[0].name])
            lead = first.shape[:first.ndim - len(self._hmc_bound.latent_sites[0].shape)]
            n = int(np.prod(lead)) if lead else 1
            if n != self._engine.C:
                raise NotImplementedError("postprocess_fn of HMCGibbs maps one state (all chains of the handle) at a time")
            flat = np.zeros((n, self._engine.D), np.float32)
            for s in self._hmc_bound.latent_sites:
                flat[:, s.z_offset:s.z_offset + s.size] = np.asarray(z[s.name], np.float32).reshape(n, s.size)
            out = {k: np.asarray(z[k]) for k in self._gibbs_sites}
            for k, v in self._constrained(flat).items():
                out[k] = v.reshape(tuple(lead) + v.shape[1:])
            return out
        return fn


class HMCECS(HMCGibbs):
    """HMC with energy-conserving subsampling (hmc_gibbs.py:502-690)."""

    def __init__(self, inner_kernel, *, num_blocks=1, proxy=None):
        super().__init__(inner_kernel, lambda *a, **k: None, None)
        if proxy is not None and not isinstance(proxy, _TaylorProxy):
            raise NotImplementedError("proxy must be HMCECS.taylor_proxy(...) or None")
        self._num_blocks, self._proxy = int(num_blocks), proxy
        self._single = True

    taylor_proxy = staticmethod(taylor_proxy)

    @property
    def default_fields(self):
        return ("z",)

    def _bind(self, model_args, model_kwargs):
        model = self.inner_kernel.model
        if not isinstance(model, families._GLM) or type(model) not in (families.LogisticRegression, families.PoissonRegression):
            raise NotImplementedError("HMCECS is implemented for LogisticRegression / PoissonRegression (plain GLM with a subsampled plate)")
        kw = dict(model_kwargs or {})
        args = list(model_args)
        m = kw.pop("subsample_size", None)
        if m is None and len(args) >= 3:
            m = args.pop(2)                                                 # model(data, labels, subsample_size) (covtype.py:66)
        if m is None:
            raise AssertionError("Cannot detect any subsample statements in the model.")      # hmc_gibbs.py:596
        bound = model.bind(*args, **kw)
        if not 0 < int(m) < bound.X.shape[0]:
            raise AssertionError("Cannot detect any subsample statements in the model.")      # size > subsample_size (:589-594)
        return bound, int(m)

    def init(self, rng_key, num_warmup, init_params=None, model_args=(), model_kwargs=None):
        """HMCECS.init :577-638 -> HMCGibbs.init :123-151 -> HMC.init (hmc.py:740-799), for one key or a batch [C, 2]."""
        keys = np.asarray(rng_key, U32)
        self._single = keys.ndim == 1
        keys = keys.reshape(-1, 2)
        C = keys.shape[0]
        bound, m = self._bind(model_args, model_kwargs)
        N, D = bound.X.shape
        self._size, self._m = N, m
        self._hmc_bound = bound
        degree = 0 if self._proxy is None else self._proxy.degree
        e = self._make_engine(keys, num_warmup, bound, dict(ecs_subsample_size=m, ecs_proxy_degree=degree, regime=_capi.REGIME_WARP))
        k2 = b2random.split_each(keys)                                       # rng_key, key_u
        rng, key_u = k2[:, 0], k2[:, 1]
        # prototype trace (seed handler, handlers.py:887-897): the latent site takes split(key_u)[1], the plate the next split
        k_plate = b2random.split_each(b2random.split_each(key_u)[:, 0])[:, 1]
        u = np.stack([_subsample_indices(k, N, m) for k in k_plate])
        if self._proxy is not None:
            self._install_proxy(e, bound)
            rng = b2random.split_each(rng)[:, 0]                             # :627-628 (gibbs_init does not use its key)
        k2 = b2random.split_each(rng)                                        # HMCGibbs.init :133: rng_key, key_z
        rng, key_z = k2[:, 0], k2[:, 1]
        e.ecs_set_indices(u)
        z0 = _flatten_init(init_params, bound, e.D) if init_params is not None else self.inner_kernel._strategy_z0(bound, e.D, C)
        e.init(key_z, int(num_warmup), z0)
        e.run(0, 0, fields=())
        hs, _, _ = self._hmc_state()
        self._u = u
        return self._pack(hs, rng, np.zeros(C, np.float32))

    def _install_proxy(self, e, bound):
        """Reference point and the full-data Taylor terms at it (contrib/ecs_proxies.py:180-186), computed once on the device."""
        ref = np.asarray(self._proxy.reference_params[bound.latent_sites[0].name], np.float32).reshape(-1)
        if ref.shape[0] != e.D:
            raise ValueError("reference_params: wrong size")
        X, y = e.X, e.y
        r = torch.from_numpy(ref).to(e.device)
        e0 = X.double() @ r.double()
        if bound.cfg["likelihood"] == _capi.LIK_BERNOULLI_LOGIT:
            s = torch.sigmoid(e0)
            ll = -(torch.clamp(e0, min=0) + torch.log1p(torch.exp(-e0.abs())) - e0 * y.double())
            d1, d2 = y.double() - s, -s * (1 - s)
        else:
            rate = torch.exp(e0)
            ll = y.double() * e0 - rate - torch.lgamma(y.double() + 1.0)
            d1, d2 = y.double() - rate, -rate
        G = (X.double().T @ d1).float().contiguous()
        H = ((X.double() * d2[:, None]).T @ X.double()).float().contiguous() if self._proxy.degree == 2 else None
        self._proxy_dev = (r.contiguous(), e0.float().contiguous(), G, H)                # borrowed by the handle: keep alive
        e.ecs_set_proxy(*self._proxy_dev, float(ll.sum().item()))

    def _pack(self, hs, rng, acc):
        sq = (lambda a: a[0]) if self._single else (lambda a: a)
        z = {"N": sq(self._u), **hs.z}
        return HMCECSState(z, hs, sq(np.asarray(rng, U32)), (), sq(acc))

    def sample(self, state, model_args=(), model_kwargs=None):
        """HMCECS.sample :640-682, all chains of the handle in lock-step."""
        e = self._engine
        if e is None:
            raise RuntimeError("sample() called before init()")
        C = e.C
        keys = np.asarray(state.rng_key, U32).reshape(C, 2)
        rng = b2random.split_each(keys)[:, 0]                                # rng_key (rng_gibbs is unused by the reference here)
        u_old = self._u
        u_new = _update_block(rng, self._num_blocks, u_old, self._size)
        hs, st, vec = self._hmc_state()
        if self.inner_kernel._dense_mass:
            self._dense = e.dense_state()
        z = vec["z"]
        pe = np.array([st[k].potential_energy for k in range(C)], np.float32)
        e.ecs_set_indices(u_new)
        pe_new_t, g_new_t = e.potential_and_grad(z)
        pe_new, g_new = pe_new_t.cpu().numpy(), g_new_t.cpu().numpy()
        with np.errstate(all="ignore"):
            acc = np.minimum(_engine.detmath(0, (pe - pe_new).astype(np.float32)), np.float32(1.0))     # clip(exp(pe - pe_new), None, 1)
        uni = np.array([_engine.prng_uniform(k, 1)[0] for k in rng], np.float32)                          # bernoulli(rng_key, accept_prob)
        take = uni < acc
        u = np.where(take[:, None], u_new, u_old).astype(np.int32)
        grad = np.where(take[:, None], g_new, vec["z_grad"]).astype(np.float32)
        pe = np.where(take, pe_new, pe).astype(np.float32)
        e.ecs_set_indices(u)
        self._u = u
        self._inner_transition(st, vec, pe, grad)
        hs, _, _ = self._hmc_state()
        return self._pack(hs, rng, acc.astype(np.float32))

    def postprocess_fn(self, args=(), kwargs=None):
        """hmc_gibbs.py:566-575: only the HMC sites are returned (the subsample indices are dropped)."""
        def fn(z):
            return {k: v for k, v in z.items() if k != "N"}
        return fn
