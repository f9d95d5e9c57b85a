"""Known answers of the reference's diagnostics tests (test/test_diagnostics.py:60-108),
checked on the oracle (direct autocovariance) and on the product's host diagnostics (FFT)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import diag as odiag


def _impls():
    out = [odiag]
    try:
        from numpyro_b200 import diagnostics as pdiag
        out.append(pdiag)
    except ImportError:
        pass
    return out


@pytest.mark.parametrize("d", _impls(), ids=lambda m: m.__name__)
def test_known_answers(d):
    x = np.arange(10.0)
    expected = np.array([1, 0.78, 0.52, 0.21, -0.13, -0.52, -0.94, -1.4, -1.91, -2.45])
    assert_allclose(d.autocorrelation(x, bias=False), expected, atol=0.01)
    assert_allclose(d.autocorrelation(x, bias=True), expected * np.arange(10, 0.0, -1) / 10, atol=0.01)
    y = np.empty((2, 10))
    y[0] = np.arange(10.0)
    y[1] = np.arange(10.0) + 1
    assert_allclose(d.gelman_rubin(y), 0.98, atol=0.01)
    z = np.random.default_rng(0).normal(size=(2, 10))
    assert_allclose(d.gelman_rubin(z.reshape(2, 2, 5).reshape(4, 5)), d.split_gelman_rubin(z))
    assert_allclose(d.effective_sample_size(np.arange(1000.0).reshape(100, 10), bias=False), 52.64, atol=0.01)


def test_product_matches_oracle_on_random_chains():
    pdiag = pytest.importorskip("numpyro_b200.diagnostics")
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(size=(4, 200, 3)), axis=1) * 0.1 + rng.normal(size=(4, 200, 3))
    assert_allclose(pdiag.effective_sample_size(x), odiag.effective_sample_size(x), rtol=1e-8)
    assert_allclose(pdiag.split_gelman_rubin(x), odiag.split_gelman_rubin(x), rtol=1e-10)
