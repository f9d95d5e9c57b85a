"""Oracle: leapfrog + iterative NUTS tree (TEST INFRASTRUCTURE ONLY).

NumPy/det-f32 restatement of numpyro/infer/hmc_util.py:
  velocity_verlet :262-311, euclidean_kinetic_energy :1183-1200, _is_turning/_momentum_angle
  :710-746, transition kernels :749-764, _combine_tree :767-848, _build_basetree :851-894,
  _double_tree :907-938, _leaf_idx_to_ckpt_idxs :941-958, _is_iterative_turning :961-981,
  _iterative_build_subtree :984-1085, build_tree :1088-1180.
One chain at a time, flat float32 vectors; the inverse mass matrix is a vector (diagonal) or a [D, D] array (dense_mass=True).  ``potential`` is any callable
``z -> (U, grad)``; in tests it is either an oracle family (fp64 rounded once) or the CUDA engine's
own potential hook (so the tree bookkeeping can be compared bit for bit).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Callable, List

import numpy as np

from . import detmath as dm
from . import prng

F = np.float32
MAX_DELTA_ENERGY = F(1000.0)        # hmc.py:188


def imm_apply(imm, r) -> np.ndarray:
    """``M^-1 r``: ``jnp.multiply`` for a diagonal, ``jnp.matmul`` for a dense inverse mass matrix (hmc_util.py:1193-1196,
    :726-731).  det-f32 order of the dense product: every row accumulates its terms left to right from +0, one rounding
    per multiplication and per addition (the engine gives row i to lane i mod 32)."""
    imm = np.asarray(imm, F)
    if imm.ndim == 1:
        return (imm * r).astype(F)
    acc = np.zeros(imm.shape[0], F)
    with np.errstate(all="ignore"):
        for j in range(imm.shape[1]):
            acc = (acc + (imm[:, j] * r[j]).astype(F)).astype(F)
    return acc


def kinetic_energy(imm, r) -> np.float32:
    """0.5 * dot(M^-1 r, r)  (hmc_util.py:1183-1200)."""
    v = imm_apply(imm, r)
    return F(F(0.5) * dm.lane_dot(v, r))


def leapfrog(potential, eps, imm, z, r, g):
    """One velocity-Verlet step (hmc_util.py:289-309); eps carries the direction sign."""
    eps = F(eps)
    half = F(F(0.5) * eps)
    r_half = (r - (half * g).astype(F)).astype(F)
    z_new = (z + (eps * imm_apply(imm, r_half)).astype(F)).astype(F)
    u_new, g_new = potential(z_new)
    r_new = (r_half - (half * g_new).astype(F)).astype(F)
    return z_new, r_new, F(u_new), np.asarray(g_new, F)


def is_turning(imm, r_left, r_right, r_sum) -> bool:
    """hmc_util.py:710-746."""
    v_left = imm_apply(imm, r_left)
    v_right = imm_apply(imm, r_right)
    mid = ((r_left + r_right).astype(F) / F(2.0)).astype(F)
    s = (r_sum - mid).astype(F)
    return bool(dm.lane_dot(v_left, s) <= 0) or bool(dm.lane_dot(v_right, s) <= 0)


def leaf_idx_to_ckpt_idxs(n: int):
    """hmc_util.py:941-958 (popcount tricks on a 32-bit leaf index)."""
    n = int(n) & 0xFFFFFFFF
    idx_max = bin(n >> 1).count("1")
    num_subtrees = bin(((~n & (n + 1)) - 1) & 0xFFFFFFFF).count("1")
    idx_min = idx_max - num_subtrees + 1
    return idx_min, idx_max


def is_iterative_turning(imm, r, r_sum, r_ckpts, r_sum_ckpts, idx_min, idx_max) -> bool:
    """hmc_util.py:961-981: scan checkpoints idx_max..idx_min, stop at first turning."""
    i = idx_max
    turning = False
    while i >= idx_min and not turning:
        sub = ((r_sum - r_sum_ckpts[i]).astype(F) + r_ckpts[i]).astype(F)
        turning = is_turning(imm, r_ckpts[i], r, sub)
        i -= 1
    return turning


@dataclass
class Tree:
    """hmc_util.TreeInfo (:36-57)."""
    z_left: np.ndarray
    r_left: np.ndarray
    g_left: np.ndarray
    z_right: np.ndarray
    r_right: np.ndarray
    g_right: np.ndarray
    z_prop: np.ndarray
    pe_prop: np.float32
    g_prop: np.ndarray
    energy_prop: np.float32
    depth: int
    weight: np.float32
    r_sum: np.ndarray
    turning: bool
    diverging: bool
    sum_accept: np.float32
    num_proposals: int


@dataclass
class TreeTrace:
    """Integer bookkeeping of one transition, for bit-exact comparison with the engine."""
    directions: List[int] = field(default_factory=list)       # going_right per doubling
    leaves: List[int] = field(default_factory=list)           # leapfrogs per doubling
    sub_turning: List[int] = field(default_factory=list)
    sub_diverging: List[int] = field(default_factory=list)
    took_new: List[int] = field(default_factory=list)         # biased-kernel accept per doubling


def _clip_max1(p):
    """jnp.clip(p, None, 1.0): NaN passes through."""
    return F(1.0) if p > F(1.0) else p


def _base_tree(potential, imm, eps, going_right, z, r, g, energy0) -> Tree:
    """hmc_util.py:851-894."""
    step = F(eps) if going_right else F(-F(eps))
    z_new, r_new, u_new, g_new = leapfrog(potential, step, imm, z, r, g)
    with np.errstate(all="ignore"):
        energy_new = F(u_new + kinetic_energy(imm, r_new))
        delta = F(energy_new - energy0)
        if np.isnan(delta):
            delta = F(np.inf)
        weight = F(-delta)
        diverging = bool(delta > MAX_DELTA_ENERGY)
        acc = _clip_max1(dm.exp(F(-delta)))
    return Tree(z_new, r_new, g_new, z_new, r_new, g_new, z_new, u_new, g_new, energy_new,
                0, weight, r_new, False, diverging, acc, 1)


def _combine(cur: Tree, new: Tree, imm, going_right: bool, key, biased: bool):
    """hmc_util.py:767-848.  Returns (tree, transition_taken)."""
    if going_right:
        zl, rl, gl = cur.z_left, cur.r_left, cur.g_left
        zr, rr, gr = new.z_right, new.r_right, new.g_right
    else:
        zl, rl, gl = new.z_left, new.r_left, new.g_left
        zr, rr, gr = cur.z_right, cur.r_right, cur.g_right
    r_sum = (cur.r_sum + new.r_sum).astype(F)
    with np.errstate(all="ignore"):
        if biased:
            p = _clip_max1(dm.exp(F(new.weight - cur.weight)))
            if new.turning or new.diverging:
                p = F(0.0)
            turning = new.turning or is_turning(imm, rl, rr, r_sum)
        else:
            p = dm.expit(F(new.weight - cur.weight))
            turning = cur.turning
    take = bool(prng.uniform(key) < p)          # random.bernoulli; NaN p -> False
    src = new if take else cur
    tree = Tree(zl, rl, gl, zr, rr, gr, src.z_prop, src.pe_prop, src.g_prop, src.energy_prop,
                cur.depth + 1, dm.logaddexp(cur.weight, new.weight), r_sum, turning,
                new.diverging, F(cur.sum_accept + new.sum_accept),
                cur.num_proposals + new.num_proposals)
    return tree, take


def _build_subtree(proto: Tree, potential, imm, eps, going_right, key, energy0,
                   r_ckpts, r_sum_ckpts) -> Tree:
    """hmc_util.py:984-1085."""
    max_n = 2 ** proto.depth
    tree = replace(proto, num_proposals=0)
    turning = False
    while tree.num_proposals < max_n and not turning and not tree.diverging:
        key, k_leaf = prng.split(key)
        if going_right:
            z, r, g = tree.z_right, tree.r_right, tree.g_right
        else:
            z, r, g = tree.z_left, tree.r_left, tree.g_left
        leaf = _base_tree(potential, imm, eps, going_right, z, r, g, energy0)
        leaf_idx = tree.num_proposals
        if leaf_idx == 0:
            new_tree = leaf
        else:
            new_tree, _ = _combine(tree, leaf, imm, going_right, k_leaf, False)
        idx_min, idx_max = leaf_idx_to_ckpt_idxs(leaf_idx)
        if leaf_idx % 2 == 0:
            r_ckpts[idx_max] = leaf.r_right
            r_sum_ckpts[idx_max] = new_tree.r_sum
        turning = is_iterative_turning(imm, leaf.r_right, new_tree.r_sum, r_ckpts, r_sum_ckpts,
                                       idx_min, idx_max)
        tree = new_tree
    return replace(tree, depth=proto.depth, turning=turning)


def build_tree(potential: Callable, imm, eps, key, z, r, pe, g, max_depth: int,
               trace: TreeTrace | None = None) -> Tree:
    """hmc_util.py:1088-1180 (``max_depth`` is the depth allowed for *this* transition)."""
    energy0 = F(pe + kinetic_energy(imm, r))
    d = z.shape[0]
    r_ckpts = np.zeros((max(max_depth, 1), d), F)
    r_sum_ckpts = np.zeros((max(max_depth, 1), d), F)
    tree = Tree(z, r, g, z, r, g, z, F(pe), g, energy0, 0, F(0.0), r, False, False, F(0.0), 0)
    while tree.depth < max_depth and not tree.turning and not tree.diverging:
        key, k_dir, k_dbl = prng.split(key, 3)
        going_right = bool(prng.bernoulli(k_dir))
        k_sub, k_fin = prng.split(k_dbl)                       # _double_tree :920
        sub = _build_subtree(tree, potential, imm, eps, going_right, k_sub, energy0,
                             r_ckpts, r_sum_ckpts)
        tree, took = _combine(tree, sub, imm, going_right, k_fin, True)
        if trace is not None:
            trace.directions.append(int(going_right))
            trace.leaves.append(int(sub.num_proposals))
            trace.sub_turning.append(int(sub.turning))
            trace.sub_diverging.append(int(sub.diverging))
            trace.took_new.append(int(took))
    return tree
