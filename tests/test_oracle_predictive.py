"""CPU checks of oracle/predictive.py (restating infer/util.py:838-1188): the per-observation log-likelihoods add up to the
likelihood part of the potential, the key plumbing follows util.py:916-918 + handlers.py:896, draws have the right law."""
import numpy as np

from oracle import families as ofam, predictive as opred, prng

F = np.float32


def test_log_likelihood_sums_to_the_likelihood_part_of_the_potential():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(400, 6))
    y = (rng.uniform(size=400) < 0.5).astype(F)
    fam = ofam.logistic_regression(X, y)
    z = rng.normal(size=6) * 0.3
    u, _ = fam.potential64(z)
    prior = np.sum(-0.5 * z * z - ofam.LOG_SQRT_2PI)
    ll = opred.log_likelihood(fam, z)
    assert ll.shape == (400,) and ll.dtype == F
    np.testing.assert_allclose(ll.astype(np.float64).sum(), -u - prior, rtol=1e-6)
    es = ofam.EightSchools(np.array([15.0, 10, 16, 11, 9, 11, 10, 18]), np.array([28.0, 8, -3, 7, -1, 1, 18, 12]))
    ze = rng.normal(size=10)
    u, _ = es.potential64(ze)
    u0, _ = ofam.EightSchools(es.sigma, es.y + 1.0).potential64(ze)          # only the likelihood term depends on y
    ll1 = opred.log_likelihood(es, ze).astype(np.float64).sum()
    ll0 = opred.log_likelihood(ofam.EightSchools(es.sigma, es.y + 1.0), ze).astype(np.float64).sum()
    np.testing.assert_allclose(ll1 - ll0, -(u - u0), rtol=1e-5)


def test_key_plumbing_and_draw_laws():
    k = prng.key(5)
    assert np.array_equal(opred.sample_keys(k, 1)[0], k)                     # one sample: the key itself
    ks = opred.sample_keys(k, 4)
    assert np.array_equal(ks, prng.split(k, 4))
    assert np.array_equal(opred.obs_key(ks[2]), prng.split(ks[2])[1])
    rng = np.random.default_rng(1)
    X = rng.normal(size=(20000, 3))
    fam = ofam.logistic_regression(X, np.zeros(20000, F))
    z = np.array([0.5, -0.25, 1.0])
    d = opred.predictive(fam, z, ks[0])
    p = 1 / (1 + np.exp(-(X @ z)))
    assert set(np.unique(d)) <= {0.0, 1.0} and abs(d.mean() - p.mean()) < 0.01
    assert np.all(opred.bernoulli_margin(fam, z, ks[0])[d == 1] >= 0)
    es = ofam.EightSchools(np.full(8, 2.0), np.zeros(8))
    ze = np.concatenate([[1.0, 0.0], np.zeros(8)])
    draws = np.stack([opred.predictive(es, ze, kk) for kk in prng.split(k, 400)])
    assert abs(draws.mean() - 1.0) < 0.15 and abs(draws.std() - 2.0) < 0.15
