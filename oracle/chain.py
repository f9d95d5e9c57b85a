"""Oracle: chain initialisation, the NUTS/HMC transition and the collection loop
(TEST INFRASTRUCTURE ONLY).

Restates numpyro/infer/hmc.py init_kernel :193-362, _hmc_next :364-414, _nuts_next :416-455,
sample_kernel :459-530, HMC.init :740-799; numpyro/infer/util.py find_valid_initial_params
:366-508 (init_to_uniform fast path :454-463); numpyro/infer/mcmc.py run :635-729 (key split per
chain :670-671); numpyro/util.py fori_collect :321-454 (collection index arithmetic).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import adapt as ad
from . import detmath as dm
from . import prng
from .tree import TreeTrace, build_tree, imm_apply, kinetic_energy, leapfrog, MAX_DELTA_ENERGY

F = np.float32


@dataclass
class HMCState:
    """numpyro.infer.hmc.HMCState (hmc.py:31-48); r / trajectory_length omitted (None for NUTS)."""
    i: int
    z: np.ndarray
    z_grad: np.ndarray
    potential_energy: np.float32
    energy: np.float32
    num_steps: int
    accept_prob: np.float32
    mean_accept_prob: np.float32
    diverging: bool
    adapt_state: ad.AdaptState
    rng_key: np.ndarray


def init_to_uniform(key, init_sites: Sequence[Tuple[str, int]], layout, potential, radius=2.0,
                    max_tries=100):
    """find_valid_initial_params with init_to_uniform (infer/util.py:417-481).

    Per attempt: (key, sub) = split(key); for each latent site in trace order draw
    uniform(sub, shape, -radius, radius) then (key, sub) = split(key).  Valid when U and grad
    are finite.  Returns (z_flat, U, grad, is_valid)."""
    offs = {n: (off, s) for n, off, s in layout}
    d = sum(s for _, s in init_sites)
    z = np.zeros(d, F)
    u, g = F(0), np.zeros(d, F)
    for _ in range(max_tries):
        key, sub = prng.split(key)
        for name, size in init_sites:
            off, s = offs[name]
            z[off:off + s] = prng.uniform(sub, size, -radius, radius)
            key, sub = prng.split(key)
        u, g = potential(z)
        if np.isfinite(u) and np.all(np.isfinite(g)):
            return z.copy(), F(u), np.asarray(g, F), True
    return z.copy(), F(u), np.asarray(g, F), False


@dataclass
class Kernel:
    """The functional ``hmc(potential_fn, algo)`` pair (hmc.py:113-191) for one chain."""
    potential: Callable
    algo: str = "NUTS"
    model_built: bool = True          # mass matrix is a one-block dict -> extra split in momentum_generator
    step_size: float = 1.0
    adapt_step_size: bool = True
    adapt_mass_matrix: bool = True
    target_accept_prob: float = 0.8
    max_tree_depth: Tuple[int, int] = (10, 10)
    find_heuristic_step_size: bool = False
    regularize_mass_matrix: bool = True
    num_steps: Optional[int] = None           # HMC only
    trajectory_length: float = 2 * np.pi      # HMC only
    dense_mass: bool = False                  # hmc.py:759-769 with dense_mass=True: one dense block over all sites

    def momentum(self, sqrt_m, key):
        """momentum_generator (hmc.py:92-110): multiply (diagonal) or dot (dense) with the unit normals."""
        if self.model_built:
            key = prng.split(key, 1)[0]
        eps = prng.normal(key, sqrt_m.shape[0])
        return imm_apply(sqrt_m, eps)

    def init(self, key, num_warmup: int, z, pe, g, inverse_mass_matrix=None) -> HMCState:
        """init_kernel (hmc.py:193-362)."""
        self.num_warmup = num_warmup
        find = None
        if self.find_heuristic_step_size:
            mk = (lambda k: prng.split(k, 1)[0]) if self.model_built else (lambda k: k)
            find = lambda step, imm, sm, z_, pe_, g_, k: ad.find_reasonable_step_size(
                self.potential, imm, sm, z_, pe_, g_, step, k, mk)
        self.adapter = ad.WarmupAdapter(num_warmup, find, self.adapt_step_size,
                                        self.adapt_mass_matrix, self.target_accept_prob,
                                        self.regularize_mass_matrix, self.dense_mass)
        k_hmc, k_wa, k_mom = prng.split(key, 3)
        wa = self.adapter.init(z, pe, g, k_wa, self.step_size, inverse_mass_matrix)
        r = self.momentum(wa.mass_matrix_sqrt, k_mom)
        energy = F(pe + kinetic_energy(wa.inverse_mass_matrix, r))
        return HMCState(0, np.asarray(z, F), np.asarray(g, F), F(pe), energy, 0, F(0), F(0),
                        False, wa, k_hmc)

    def _hmc_next(self, eps, imm, z, r, pe, g, key):
        """_hmc_next (hmc.py:364-414)."""
        with np.errstate(all="ignore"):
            if self.num_steps is not None:
                n = self.num_steps
            else:
                n = int(np.ceil(F(F(self.trajectory_length) / eps)))
                eps = F(F(self.trajectory_length) / F(n))
            z1, r1, pe1, g1 = z, r, pe, g
            for _ in range(n):
                z1, r1, pe1, g1 = leapfrog(self.potential, eps, imm, z1, r1, g1)
            e_old = F(pe + kinetic_energy(imm, r))
            e_new = F(pe1 + kinetic_energy(imm, r1))
            delta = F(e_new - e_old)
            if np.isnan(delta):
                delta = F(np.inf)
            acc = dm.exp(F(-delta))
            acc = F(1.0) if acc > F(1.0) else acc
            diverging = bool(delta > MAX_DELTA_ENERGY)
            take = bool(prng.uniform(key) < acc)
        if take:
            return z1, pe1, g1, e_new, n, acc, diverging
        return z, pe, g, e_old, n, acc, diverging

    def sample(self, s: HMCState, trace: Optional[TreeTrace] = None) -> HMCState:
        """sample_kernel (hmc.py:459-530)."""
        key, k_mom, k_tr = prng.split(s.rng_key, 3)
        a = s.adapt_state
        r = self.momentum(a.mass_matrix_sqrt, k_mom)
        in_warmup = s.i < self.num_warmup
        if self.algo == "NUTS":
            depth = self.max_tree_depth[0] if in_warmup else self.max_tree_depth[1]
            t = build_tree(self.potential, a.inverse_mass_matrix, a.step_size, k_tr, s.z, r,
                           s.potential_energy, s.z_grad, depth, trace)
            with np.errstate(all="ignore"):
                acc = F(t.sum_accept / F(t.num_proposals))
            z, pe, g, energy, n, div = t.z_prop, t.pe_prop, t.g_prop, t.energy_prop, t.num_proposals, t.diverging
        else:
            z, pe, g, energy, n, acc, div = self._hmc_next(a.step_size, a.inverse_mass_matrix, s.z, r,
                                                           s.potential_energy, s.z_grad, k_tr)
        if in_warmup:
            a = self.adapter.update(s.i, acc, z, pe, g, a)
        itr = s.i + 1
        n_mean = itr if in_warmup else itr - self.num_warmup
        with np.errstate(all="ignore"):
            mean_acc = F(s.mean_accept_prob + F(F(acc - s.mean_accept_prob) / F(n_mean)))
        return HMCState(itr, z, g, F(pe), F(energy), int(n), F(acc), mean_acc, bool(div), a, key)


def chain_keys(key, num_chains: int) -> np.ndarray:
    """mcmc.py:670-671: one chain uses the user key itself, C>1 uses split(key, C)."""
    key = np.asarray(key, np.uint32)
    if key.ndim == 2:
        return key
    return key[None] if num_chains == 1 else prng.split(key, num_chains)


def run_chain(kernel: Kernel, family, chain_key, num_warmup: int, num_samples: int,
              thinning: int = 1, init_z=None, collect_warmup=False,
              fields: Sequence[str] = ("z", "diverging", "num_steps", "accept_prob",
                                       "potential_energy", "energy", "step_size"),
              traces: Optional[List[TreeTrace]] = None, inverse_mass_matrix=None):
    """One chain of ``MCMC.run`` (mcmc.py:466-521) = HMC.init + fori_collect.

    Returns (dict of per-sample arrays, last HMCState).  Collection follows util.py:368-403:
    iterations ``lower..upper`` with ``lower = 0 if collect_warmup else num_warmup``, a sample is
    written at ``(i - start) // thinning`` where ``start = lower + (upper-lower) % thinning``."""
    rng_key, k_init = prng.split(chain_key)                        # hmc.py:744-750
    pot = kernel.potential
    if init_z is None:
        z, pe, g, ok = init_to_uniform(k_init, family.init_sites, family.layout, pot)
        if not ok:
            raise RuntimeError("Cannot find valid initial parameters. Please check your model again.")
    else:
        z = np.asarray(init_z, F)
        pe, g = pot(z)
    state = kernel.init(rng_key, num_warmup, z, pe, g, inverse_mass_matrix)
    upper = num_warmup + num_samples
    lower = 0 if collect_warmup else num_warmup
    size = (upper - lower) // thinning
    start = lower + (upper - lower) % thinning
    out: Dict[str, list] = {f: [] for f in fields}
    for i in range(upper):
        tr = TreeTrace() if traces is not None else None
        state = kernel.sample(state, tr)
        if traces is not None:
            traces.append(tr)
        if i >= start and ((i - start) + 1) % thinning == 0:
            rec = {"z": state.z, "diverging": state.diverging, "num_steps": state.num_steps,
                   "accept_prob": state.accept_prob, "potential_energy": state.potential_energy,
                   "energy": state.energy, "step_size": state.adapt_state.step_size,
                   "mean_accept_prob": state.mean_accept_prob}
            for f in fields:
                out[f].append(np.copy(rec[f]))
    res = {f: np.asarray(v) for f, v in out.items()}
    assert all(v.shape[0] == size for v in res.values())
    return res, state
