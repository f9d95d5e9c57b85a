"""Key helpers with jax.random's key layout (uint32[2]); the streams are produced on the device
(csrc/prng.cuh) through the C ABI's PRNG entry points."""
from __future__ import annotations

import numpy as np

from . import engine as _engine


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` / ``key_data(jax.random.key(seed))``: (hi32, lo32) of the seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


key = PRNGKey


def split(key, num: int = 2) -> np.ndarray:
    """``jax.random.split`` (partitionable Threefry), computed by the engine."""
    return _engine.prng_split(np.asarray(key, np.uint32).reshape(1, 2), int(num))[0]


def split_each(keys, num: int = 2) -> np.ndarray:
    """``vmap(lambda k: jax.random.split(k, num))(keys)``: ``[n, 2]`` keys -> ``[n, num, 2]``, one device call."""
    return _engine.prng_split(np.asarray(keys, np.uint32).reshape(-1, 2), int(num))


def is_prng_key(k) -> bool:
    """numpyro.util.is_prng_key (util.py:176-182) for raw uint32[2] key data."""
    k = np.asarray(k)
    return k.dtype == np.uint32 and k.shape == (2,)
