"""Accuracy of the det-f32 transcendental algorithms (oracle/detmath.py) against float64."""
import numpy as np
from scipy.special import erfinv as sp_erfinv

from oracle import detmath as dm

F = np.float32


def _max_ulp(f, ref, xs):
    worst = 0.0
    for x in xs:
        r = float(ref(np.float64(F(x))))
        if not np.isfinite(r) or r == 0:
            continue
        worst = max(worst, abs(float(f(F(x))) - r) / float(abs(np.spacing(F(r)))))
    return worst


def test_exp_log_log1p_within_2ulp():
    rng = np.random.default_rng(0)
    assert _max_ulp(dm.exp, np.exp, np.concatenate([rng.uniform(-100, 88, 4000), rng.uniform(-1, 1, 2000)])) < 1.5
    assert _max_ulp(dm.log, np.log, np.concatenate([np.exp(rng.uniform(-100, 88, 4000)), rng.uniform(0.5, 2, 2000)])) < 1.5
    assert _max_ulp(dm.log1p, np.log1p, np.concatenate([rng.uniform(-0.999, 5, 4000), rng.uniform(-1e-3, 1e-3, 2000)])) < 2.5


def test_special_values():
    assert dm.exp(F(0)) == 1 and dm.exp(F(-np.inf)) == 0 and dm.exp(F(np.inf)) == np.inf
    assert dm.exp(F(89)) == np.inf and dm.exp(F(-104)) == 0 and np.isnan(dm.exp(F(np.nan)))
    assert dm.exp(F(-100)) == np.exp(F(-100))                  # sub-normal result
    assert dm.log(F(1)) == 0 and dm.log(F(0)) == -np.inf and np.isnan(dm.log(F(-1)))
    assert dm.log(F(np.inf)) == np.inf
    assert dm.log1p(F(-1)) == -np.inf and dm.log1p(F(0)) == 0 and dm.log1p(F(1e-10)) == F(1e-10)
    assert dm.expit(F(0)) == F(0.5) and dm.expit(F(-np.inf)) == 0 and dm.expit(F(np.inf)) == 1
    assert dm.logaddexp(F(-np.inf), F(-np.inf)) == -np.inf
    assert dm.logaddexp(F(0), F(-np.inf)) == 0
    assert abs(dm.logaddexp(F(1), F(2)) - np.logaddexp(1.0, 2.0)) < 1e-6
    assert dm.erfinv(F(1)) == np.inf and dm.erfinv(F(-1)) == -np.inf and dm.erfinv(F(0)) == 0


def test_erfinv_close_to_scipy():
    xs = np.random.default_rng(1).uniform(-0.999, 0.999, 3000)
    got = np.array([dm.erfinv(F(x)) for x in xs], np.float64)
    np.testing.assert_allclose(got, sp_erfinv(np.float64(F(xs))), rtol=5e-6, atol=1e-7)


def test_lane_sum_order():
    x = np.random.default_rng(2).normal(size=70).astype(F)
    p = np.zeros(32, F)
    for d in range(70):
        p[d % 32] = p[d % 32] + x[d]
    for off in (16, 8, 4, 2, 1):
        p = (p + p[np.arange(32) ^ off]).astype(F)
    assert dm.lane_sum(x) == p[0]
    assert abs(float(dm.lane_sum(x)) - float(x.astype(np.float64).sum())) < 1e-4
