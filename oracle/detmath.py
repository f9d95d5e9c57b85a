"""Oracle: the "det-f32" numeric convention (TEST INFRASTRUCTURE ONLY).

The reference leaves exp/log/log1p/erfinv and reduction order to XLA (an un-vendored
dependency whose CPU and GPU lowerings differ from each other in the last bits).  To make the
integer bookkeeping of the NUTS tree bit-reproducible between this oracle and the CUDA engine,
both sides implement the same explicit binary32 algorithms, written with one IEEE rounding per
operation and no FMA contraction:

* ``exp``  -- Cody-Waite range reduction + degree-5 polynomial (Cephes ``expf`` constants),
  two-step power-of-two scaling so sub-normal results round once.
* ``log``  -- frexp + degree-8 polynomial (Cephes ``logf`` constants).
* ``log1p``-- ``u = 1 + x;  log(u) * x / (u - 1)`` (Kahan), exact for ``u == 1``.
* ``erfinv`` -- Giles' two-branch degree-8 polynomial on ``w = -log1p(-x*x)``, the algorithm XLA
  uses for f32 ``ErfInv`` (constants as recalled in SURVEY.md App. B.3; pinned through the two
  ``jax.random.normal`` values printed in JAX's documentation).
* ``lane_sum`` -- reduction order of a D-vector: 32 lane partials ``p[l] = sum_k x[l+32k]``
  (k ascending), then an xor-butterfly ``p[l] += p[l^off]`` for off = 16, 8, 4, 2, 1.

Every function takes/returns ``np.float32`` scalars; max error vs the correctly rounded result is
about 2 ulp (tests/test_oracle_detmath.py), i.e. within the variation between XLA back ends.
The device twin is numpyro_b200/csrc/detmath.cuh.
"""
from __future__ import annotations

import numpy as np

F = np.float32
_I32 = np.int32
_U32 = np.uint32


def _bits(x) -> int:
    return int(np.array(x, dtype=F).view(_U32))


def _from_bits(b: int) -> np.float32:
    return np.array(b & 0xFFFFFFFF, dtype=_U32).view(F)[()]


def _pow2(k: int) -> np.float32:
    """2**k for -126 <= k <= 127 as an exact float32."""
    return _from_bits((k + 127) << 23)


_LOG2E = F(1.44269504088896341)
_LN2_HI = F(0.693359375)
_LN2_LO = F(-2.12194440e-4)
_EXP_P = tuple(F(c) for c in (1.9875691500e-4, 1.3981999507e-3, 8.3334519073e-3,
                              4.1665795894e-2, 1.6666665459e-1, 5.0000001201e-1))
_EXP_HI = F(88.72283935546875)    # largest x with finite expf
_EXP_LO = F(-103.972084045410)     # below this expf underflows to +0


def exp(x) -> np.float32:
    x = F(x)
    if np.isnan(x):
        return x
    if x > _EXP_HI:
        return F(np.inf)
    if x < _EXP_LO:
        return F(0.0)
    with np.errstate(all="ignore"):
        t = F(x * _LOG2E)
        half = F(0.5) if t >= 0 else F(-0.5)
        k = int(np.trunc(F(t + half)))
        kf = F(k)
        r = F(F(x - F(kf * _LN2_HI)) - F(kf * _LN2_LO))
        p = _EXP_P[0]
        for c in _EXP_P[1:]:
            p = F(F(p * r) + c)
        z = F(r * r)
        y = F(F(F(p * z) + r) + F(1.0))
        k1 = int(k / 2)          # truncation toward zero
        k2 = k - k1
        return F(F(y * _pow2(k1)) * _pow2(k2))


_SQRTHF = F(0.707106781186547524)
_LOG_P = tuple(F(c) for c in (7.0376836292e-2, -1.1514610310e-1, 1.1676998740e-1,
                              -1.2420140846e-1, 1.4249322787e-1, -1.6668057665e-1,
                              2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1))
_TWO23 = F(8388608.0)
_FLT_MIN = F(1.17549435e-38)


def log(x) -> np.float32:
    x = F(x)
    if np.isnan(x) or x < 0:
        return F(np.nan)
    if x == 0:
        return F(-np.inf)
    if np.isinf(x):
        return x
    with np.errstate(all="ignore"):
        e = 0
        if x < _FLT_MIN:
            x = F(x * _TWO23)
            e = -23
        b = _bits(x)
        e += ((b >> 23) & 0xFF) - 126
        m = _from_bits((b & 0x007FFFFF) | 0x3F000000)     # mantissa in [0.5, 1)
        if m < _SQRTHF:
            e -= 1
            m = F(F(m + m) - F(1.0))
        else:
            m = F(m - F(1.0))
        ef = F(e)
        z = F(m * m)
        p = _LOG_P[0]
        for c in _LOG_P[1:]:
            p = F(F(p * m) + c)
        y = F(F(p * m) * z)
        y = F(y + F(_LN2_LO * ef))
        y = F(y - F(F(0.5) * z))
        r = F(m + y)
        return F(r + F(_LN2_HI * ef))


def log1p(x) -> np.float32:
    x = F(x)
    if np.isnan(x):
        return x
    with np.errstate(all="ignore"):
        u = F(F(1.0) + x)
        if u == F(1.0):
            return x
        if np.isinf(u):
            return u
        return F(F(log(u) * x) / F(u - F(1.0)))


def expit(x) -> np.float32:
    """jax.scipy.special.expit = lax.logistic = 1 / (1 + exp(-x))."""
    with np.errstate(all="ignore"):
        return F(F(1.0) / F(F(1.0) + exp(F(-F(x)))))


def logaddexp(a, b) -> np.float32:
    """jnp.logaddexp: amax + log1p(exp(-|a-b|)); a+b when a-b is NaN (both -inf / +inf)."""
    a = F(a)
    b = F(b)
    with np.errstate(all="ignore"):
        d = F(a - b)
        if np.isnan(d):
            return F(a + b)
        amax = a if a >= b else b
        return F(amax + log1p(exp(F(-abs(d)))))


def powf(x, y) -> np.float32:
    """x**y for x > 0 as exp(y*log(x)) (used for t**-0.75 in dual averaging)."""
    return exp(F(F(y) * log(F(x))))


_ERFINV_A = tuple(F(c) for c in (2.81022636e-08, 3.43273939e-07, -3.5233877e-06,
                                 -4.39150654e-06, 0.00021858087, -0.00125372503,
                                 -0.00417768164, 0.246640727, 1.50140941))
_ERFINV_B = tuple(F(c) for c in (-0.000200214257, 0.000100950558, 0.00134934322,
                                 -0.00367342844, 0.00573950773, -0.0076224613,
                                 0.00943887047, 1.00167406, 2.83297682))


def erfinv(x) -> np.float32:
    x = F(x)
    with np.errstate(all="ignore"):
        if abs(x) == F(1.0):
            return F(np.copysign(np.inf, x))
        w = F(-log1p(F(-F(x * x))))
        if w < F(5.0):
            w = F(w - F(2.5))
            cs = _ERFINV_A
        else:
            w = F(F(np.sqrt(w)) - F(3.0))
            cs = _ERFINV_B
        p = cs[0]
        for c in cs[1:]:
            p = F(c + F(p * w))
        return F(p * x)


_BUTTERFLY = [np.arange(32) ^ off for off in (16, 8, 4, 2, 1)]


def lane_sum(x) -> np.float32:
    """Canonical reduction order of a D-vector (see module docstring)."""
    x = np.asarray(x, dtype=F).ravel()
    d = x.shape[0]
    pad = (-d) % 32
    if pad:
        x = np.concatenate([x, np.zeros(pad, F)])
    rows = x.reshape(-1, 32)
    p = np.zeros(32, F)
    for k in range(rows.shape[0]):
        p = (p + rows[k]).astype(F)
    for perm in _BUTTERFLY:
        p = (p + p[perm]).astype(F)
    return p[0]


def lane_dot(a, b) -> np.float32:
    a = np.asarray(a, F)
    b = np.asarray(b, F)
    return lane_sum((a * b).astype(F))
