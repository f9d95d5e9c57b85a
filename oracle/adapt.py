"""Oracle: warm-up adaptation (TEST INFRASTRUCTURE ONLY).

NumPy/det-f32 restatement of numpyro/infer/hmc_util.py: dual_averaging :60-130,
welford_covariance :133-239 (diagonal and dense), find_reasonable_step_size :314-384,
build_adaptation_schedule :387-436, warmup_adapter :518-707.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import detmath as dm
from . import prng
from .tree import imm_apply, kinetic_energy, leapfrog

F = np.float32
TINY = np.finfo(np.float32).tiny
FMAX = np.finfo(np.float32).max
LOG10 = dm.log(F(10.0))


def build_adaptation_schedule(num_steps: int) -> List[Tuple[int, int]]:
    """hmc_util.py:387-436 (Stan's windowed schedule); returns inclusive (start, end) pairs."""
    if num_steps < 20:
        return [(0, num_steps - 1)]
    start_buf, end_buf, init_win = 75, 50, 25
    if start_buf + end_buf + init_win > num_steps:
        start_buf = int(0.15 * num_steps)
        end_buf = int(0.1 * num_steps)
        init_win = num_steps - start_buf - end_buf
    sched = [(0, start_buf - 1)]
    end_start = num_steps - end_buf
    next_size, next_start = init_win, start_buf
    while next_start < end_start:
        cur_start, cur_size = next_start, next_size
        if 3 * cur_size <= end_start - cur_start:
            next_size = 2 * cur_size
        else:
            cur_size = end_start - cur_start
        next_start = cur_start + cur_size
        sched.append((cur_start, next_start - 1))
    sched.append((end_start, num_steps - 1))
    return sched


# ---------------------------------------------------------------- dual averaging (:60-130)
@dataclass
class DAState:
    x_t: np.float32
    x_avg: np.float32
    g_avg: np.float32
    t: int
    prox: np.float32


def da_init(prox_center) -> DAState:
    return DAState(F(0.0), F(0.0), F(0.0), 0, F(prox_center))


def da_update(g, s: DAState, t0=10, kappa=0.75, gamma=0.05) -> DAState:
    with np.errstate(all="ignore"):
        t = s.t + 1
        inv = F(F(1.0) / F(t + t0))
        g_avg = F(F(F(F(1.0) - inv) * s.g_avg) + F(F(g) / F(t + t0)))
        x_t = F(s.prox - F(F(F(np.sqrt(F(t))) / F(gamma)) * g_avg))
        w = dm.powf(F(t), F(-kappa))
        x_avg = F(F(F(F(1.0) - w) * s.x_avg) + F(w * x_t))
    return DAState(x_t, x_avg, g_avg, t, s.prox)


# ---------------------------------------------------------------- Welford (:133-239)
@dataclass
class WelfordState:
    mean: np.ndarray
    m2: np.ndarray
    n: int


def welford_init(d: int, dense: bool = False) -> WelfordState:
    return WelfordState(np.zeros(d, F), np.zeros((d, d) if dense else d, F), 0)


def welford_update(z, s: WelfordState) -> WelfordState:
    n = s.n + 1
    delta_pre = (z - s.mean).astype(F)
    mean = (s.mean + (delta_pre / F(n)).astype(F)).astype(F)
    delta_post = (z - mean).astype(F)
    if s.m2.ndim == 1:
        m2 = (s.m2 + (delta_pre * delta_post).astype(F)).astype(F)
    else:                                                          # :193-194  m2 + outer(delta_post, delta_pre)
        m2 = (s.m2 + np.outer(delta_post, delta_pre).astype(F)).astype(F)
    return WelfordState(mean, m2, n)


def cholesky_lower(a) -> np.ndarray:
    """``jnp.linalg.cholesky(a)`` (lower factor; jax symmetrises its input by default: (a + a^T) / 2) in det-f32: element
    (i, j), j <= i:  s = a_sym[i, j] - sum_{k < j} L[i, k] * L[j, k]  (k ascending, one rounding per operation), then
    sqrt(s) on the diagonal, s / L[j, j] below it.  A non-positive pivot gives NaN, as LAPACK / XLA do."""
    a = np.asarray(a, F)
    d = a.shape[0]
    with np.errstate(all="ignore"):
        sym = ((a + a.T).astype(F) / F(2.0)).astype(F)
        L = np.zeros((d, d), F)
        for j in range(d):
            s = sym[j:, j].copy()
            for k in range(j):
                s = (s - (L[j:, k] * L[j, k]).astype(F)).astype(F)
            piv = np.sqrt(s[0]).astype(F) if s[0] > 0 else F(np.nan)
            L[j, j] = piv
            L[j + 1:, j] = (s[1:] / piv).astype(F)
    return L


def solve_lower_identity(t) -> np.ndarray:
    """``solve_triangular(t, identity, lower=True)`` = t^-1 by forward substitution, column by column:
    x_i = (e_i - sum_{k < i} t[i, k] x_k) / t[i, i], k ascending, one rounding per operation."""
    t = np.asarray(t, F)
    d = t.shape[0]
    x = np.zeros((d, d), F)
    with np.errstate(all="ignore"):
        for i in range(d):
            s = np.zeros(d, F)
            s[i] = F(1.0)
            for k in range(i):
                s = (s - (t[i, k] * x[k, :]).astype(F)).astype(F)
            x[i, :] = (s / t[i, i]).astype(F)
    return x


def mass_matrix_roots(imm):
    """hmc_util.py:228-233 / :499-509: (mass_matrix_sqrt, mass_matrix_sqrt_inv) of a dense inverse mass matrix:
    tril_inv = swapaxes(cholesky(imm[::-1, ::-1])[::-1, ::-1]);  sqrt = solve_triangular(tril_inv, I, lower=True)."""
    lc = cholesky_lower(np.asarray(imm, F)[::-1, ::-1])
    tril_inv = np.ascontiguousarray(lc[::-1, ::-1].T)
    return solve_lower_identity(tril_inv), tril_inv


def welford_final(s: WelfordState, regularize: bool):
    """Returns (inverse_mass_matrix, mass_matrix_sqrt, mass_matrix_sqrt_inv)."""
    with np.errstate(all="ignore"):
        cov = (s.m2 / F(s.n - 1)).astype(F)
        if regularize:
            scaled = (F(F(s.n) / F(s.n + 5)) * cov).astype(F)
            shrink = F(F(1e-3) * F(F(5.0) / F(s.n + 5)))
            if cov.ndim == 1:
                cov = (scaled + shrink).astype(F)
            else:                                   # :222  scaled_cov + shrinkage * identity
                cov = (scaled + (shrink * np.identity(cov.shape[0], dtype=F)).astype(F)).astype(F)
        if cov.ndim == 2:
            sqrt_m, sqrt_inv = mass_matrix_roots(cov)
            return cov, sqrt_m, sqrt_inv
        sqrt_inv = np.sqrt(cov).astype(F)          # mass_matrix_sqrt_inv  (tril_inv)
        sqrt_m = (F(1.0) / sqrt_inv).astype(F)     # mass_matrix_sqrt      (cov_inv_sqrt)
    return cov, sqrt_m, sqrt_inv


# ---------------------------------------------------------------- step-size heuristic (:314-384)
def find_reasonable_step_size(potential: Callable, imm, sqrt_m, z, pe, g, init_step, key,
                              momentum_key_fn: Callable = lambda k: k):
    """``momentum_key_fn`` maps the momentum key to the key actually fed to ``normal`` (a
    model-built kernel splits once more because its mass matrix is a one-block dict)."""
    target = dm.log(F(0.8))
    step = F(init_step)
    last_dir, direction = 0, 0
    d = z.shape[0]
    with np.errstate(all="ignore"):
        while True:
            not_small = (step > TINY) or (direction >= 0)
            not_large = (step < FMAX) or (direction <= 0)
            if not (not_small and not_large and (last_dir == 0 or direction == last_dir)):
                break
            key, k_mom = prng.split(key)
            step = F(F(2.0) ** direction * step)
            # NB: the reference passes inverse_mass_matrix as momentum_generator's mass_matrix_sqrt
            # argument here (hmc_util.py:355), so r = M^-1 * eps.
            r = imm_apply(imm, prng.normal(momentum_key_fn(k_mom), d))
            _, r_new, pe_new, _ = leapfrog(potential, step, imm, z, r, g)
            e_cur = F(kinetic_energy(imm, r) + pe)
            e_new = F(kinetic_energy(imm, r_new) + pe_new)
            delta = F(e_new - e_cur)
            new_dir = 1 if target < F(-delta) else -1
            last_dir, direction = direction, new_dir
    return step


# ---------------------------------------------------------------- warm-up adapter (:518-707)
@dataclass
class AdaptState:
    """hmc_util.HMCAdaptState (:18-30)."""
    step_size: np.float32
    inverse_mass_matrix: np.ndarray
    mass_matrix_sqrt: np.ndarray
    mass_matrix_sqrt_inv: np.ndarray
    ss_state: DAState
    mm_state: WelfordState
    window_idx: int
    rng_key: np.ndarray


@dataclass
class WarmupAdapter:
    num_adapt_steps: int
    find_step: Optional[Callable] = None      # (step, imm, sqrt_m, z, pe, g, key) -> step
    adapt_step_size: bool = True
    adapt_mass_matrix: bool = True
    target_accept_prob: float = 0.8
    regularize_mass_matrix: bool = True
    dense_mass: bool = False

    def __post_init__(self):
        self.schedule = build_adaptation_schedule(self.num_adapt_steps)
        self.num_windows = len(self.schedule)

    def init(self, z, pe, g, key, step_size=1.0, inverse_mass_matrix=None) -> AdaptState:
        key, k_ss = prng.split(key)
        d = z.shape[0]
        if inverse_mass_matrix is None:                       # _initialize_mass_matrix :487-493
            imm = np.identity(d, dtype=F) if self.dense_mass else np.ones(d, F)
            sqrt_m = sqrt_inv = imm
        elif self.dense_mass:                                 # :495-509
            imm = np.asarray(inverse_mass_matrix, F)
            if imm.ndim == 1:
                imm = np.diag(imm).astype(F)
            sqrt_m, sqrt_inv = mass_matrix_roots(imm)
        else:
            imm = np.asarray(inverse_mass_matrix, F)
            if imm.ndim == 2:
                imm = np.ascontiguousarray(np.diag(imm))
            sqrt_inv = np.sqrt(imm).astype(F)
            sqrt_m = (F(1.0) / sqrt_inv).astype(F)
        step = F(step_size)
        if self.adapt_step_size and self.find_step is not None:
            step = self.find_step(step, imm, sqrt_m, z, pe, g, k_ss)
        with np.errstate(all="ignore"):
            ss = da_init(dm.log(F(F(10.0) * step)))
        return AdaptState(step, imm, sqrt_m, sqrt_inv, ss, welford_init(d, self.dense_mass), 0, key)

    def update(self, t: int, accept_prob, z, pe, g, s: AdaptState) -> AdaptState:
        key, k_ss = prng.split(s.rng_key)
        step, ss = s.step_size, s.ss_state
        with np.errstate(all="ignore"):
            if self.adapt_step_size:
                ss = da_update(F(F(self.target_accept_prob) - F(accept_prob)), ss)
                log_step = ss.x_avg if t == self.num_adapt_steps - 1 else ss.x_t
                step = dm.exp(log_step)
                step = F(min(max(step, TINY), FMAX))       # jnp.clip (NaN propagates via max/min)
                if np.isnan(log_step):
                    step = F(np.nan)
        middle = 0 < s.window_idx < self.num_windows - 1
        mm = s.mm_state
        if self.adapt_mass_matrix and middle:
            mm = welford_update(z, mm)
        at_end = t == self.schedule[min(s.window_idx, self.num_windows - 1)][1]
        widx = s.window_idx + 1 if at_end else s.window_idx
        out = AdaptState(step, s.inverse_mass_matrix, s.mass_matrix_sqrt, s.mass_matrix_sqrt_inv,
                         ss, mm, widx, key)
        if at_end and middle:
            imm, sqrt_m, sqrt_inv = out.inverse_mass_matrix, out.mass_matrix_sqrt, out.mass_matrix_sqrt_inv
            if self.adapt_mass_matrix:
                imm, sqrt_m, sqrt_inv = welford_final(mm, self.regularize_mass_matrix)
                mm = welford_init(z.shape[0], self.dense_mass)
            if self.adapt_step_size:
                if self.find_step is not None:
                    step = self.find_step(step, imm, sqrt_m, z, pe, g, k_ss)
                with np.errstate(all="ignore"):
                    ss = da_init(F(LOG10 + dm.log(step)))
            out = replace(out, step_size=step, inverse_mass_matrix=imm, mass_matrix_sqrt=sqrt_m,
                          mass_matrix_sqrt_inv=sqrt_inv, ss_state=ss, mm_state=mm)
        return out
