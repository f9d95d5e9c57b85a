"""Oracle: JAX's Threefry-2x32 PRNG, restated in NumPy (TEST INFRASTRUCTURE ONLY).

The reference (numpyro 0.21.0) draws every random number through ``jax.random`` -- call sites
numpyro/infer/mcmc.py:671, hmc.py:94,102,335,472-474,745-750, hmc_util.py:355,565,656,804,920,
1005,1161-1162, infer/util.py:423,460-463.  The generator itself lives in the un-vendored
dependency ``jax>=0.7`` (pyproject.toml:22-28), which is absent from this image, so this file
restates the published algorithm:

* Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
  Random123 v1.09 known-answer file ``kat_vectors``).
* JAX key derivation in *partitionable* mode (``jax_threefry_partitionable=True``, the default
  since jax 0.5 and therefore for the reference's ``jax>=0.7``): ``split(key, n)[i]`` is the
  2-word Threefry output for counter ``(0, i)``; ``random_bits(key, shape)[i]`` is ``x0 ^ x1`` of
  the output for the 64-bit counter ``i`` split as ``(hi, lo)``.
* ``uniform``: mantissa trick ``bitcast((bits >> 9) | 0x3F800000) - 1``; ``normal``:
  ``sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))``; ``bernoulli``: ``uniform < p``.

Pins: tests/test_oracle_prng.py checks the Random123 KATs and the values printed in JAX's docs.
"""
from __future__ import annotations

import numpy as np

from . import detmath as dm

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32-20 block function on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, U32)
        k1 = np.asarray(k1, U32)
        ks = (k0, k1, (k0 ^ k1 ^ U32(0x1BD11BDA)).astype(U32))
        x0 = (np.asarray(c0, U32) + ks[0]).astype(U32)
        x1 = (np.asarray(c1, U32) + ks[1]).astype(U32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(U32)
            j = g + 1
            x0 = (x0 + ks[j % 3]).astype(U32)
            x1 = (x1 + ks[(j + 1) % 3] + U32(j)).astype(U32)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` / ``PRNGKey(seed)`` key data: (hi32, lo32) of the 64-bit seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def split(k, n: int = 2) -> np.ndarray:
    """``jax.random.split(key, n)`` (partitionable): row i = threefry(key; (0, i))."""
    k = np.asarray(k, U32)
    i = np.arange(n, dtype=U32)
    x0, x1 = threefry2x32(k[0], k[1], np.zeros(n, U32), i)
    return np.stack([x0, x1], axis=-1)


def split_legacy(k, n: int = 2) -> np.ndarray:
    """Pre-0.5 (non-partitionable) split, kept only to reproduce the legacy doc vector."""
    k = np.asarray(k, U32)
    cnt = np.arange(2 * n, dtype=U32)
    x0, x1 = threefry2x32(k[0], k[1], cnt[:n], cnt[n:])
    return np.concatenate([x0, x1]).reshape(n, 2)


def random_bits(k, n=None) -> np.ndarray:
    """32 random bits for a flat shape ``(n,)`` (``n=None`` -> scalar shape ``()``)."""
    k = np.asarray(k, U32)
    m = 1 if n is None else int(n)
    idx = np.arange(m, dtype=np.uint64)
    hi = (idx >> np.uint64(32)).astype(U32)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(U32)
    x0, x1 = threefry2x32(k[0], k[1], hi, lo)
    bits = (x0 ^ x1).astype(U32)
    return bits[0] if n is None else bits


def _bits_to_unit(bits) -> np.ndarray:
    f = ((np.asarray(bits, U32) >> U32(9)) | U32(0x3F800000)).astype(U32).view(np.float32)
    return (f - np.float32(1.0)).astype(np.float32)


def uniform(k, n=None, lo=0.0, hi=1.0) -> np.ndarray:
    """``jax.random.uniform(key, shape, float32, lo, hi)``: max(lo, u*(hi-lo)+lo)."""
    lo = np.float32(lo)
    hi = np.float32(hi)
    u = _bits_to_unit(random_bits(k, n))
    scale = np.float32(hi - lo)
    v = (u * scale).astype(np.float32) + lo
    return np.maximum(lo, v.astype(np.float32)).astype(np.float32)


def bernoulli(k, p=0.5, n=None):
    """``jax.random.bernoulli(key, p)``: uniform(key) < p."""
    return uniform(k, n) < np.float32(p)


_NEG1_PLUS = np.nextafter(np.float32(-1.0), np.float32(0.0))
_SQRT2 = np.float32(np.sqrt(2.0))


def normal(k, n=None) -> np.ndarray:
    """``jax.random.normal(key, shape, float32)`` = sqrt(2) * erfinv(uniform(-1+ulp, 1))."""
    u = uniform(k, n, _NEG1_PLUS, 1.0)
    if n is None:
        return np.float32(_SQRT2 * dm.erfinv(u))
    return np.array([_SQRT2 * dm.erfinv(x) for x in u], dtype=np.float32)
