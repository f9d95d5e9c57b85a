"""CPU checks of oracle/ecs.py (restating numpyro/infer/hmc_gibbs.py:502-690 + contrib/ecs_proxies.py): the analytic gradient of
the bias-corrected estimator against finite differences, exactness of the Taylor proxy at full sample size, the block update's
index arithmetic, and the integer draws."""
import numpy as np

from oracle import ecs, prng

F = np.float32


def _data(n=400, d=5, lik="bernoulli", seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)).astype(F) * (0.3 if lik == "poisson" else 1.0)
    beta = rng.normal(size=d) * 0.5
    eta = X @ beta
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-eta))).astype(F) if lik == "bernoulli" else rng.poisson(np.exp(eta)).astype(F)
    return X, y, beta


def test_gradient_matches_finite_differences_and_estimator_is_consistent():
    for lik in ("bernoulli", "poisson"):
        X, y, beta = _data(lik=lik)
        ref = beta + 0.05
        idx = np.random.default_rng(1).integers(0, 400, size=60).astype(np.int32)
        for degree in (None, 1, 2):
            proxy = None if degree is None else ecs.TaylorProxy.build(X, y, lik, ref, degree)
            z = beta + np.random.default_rng(2).normal(size=5) * 0.1
            u, g = ecs.ecs_potential64(X, y, lik, z, idx, proxy)
            for j in range(5):
                dz = np.zeros(5); dz[j] = 1e-6
                up, _ = ecs.ecs_potential64(X, y, lik, z + dz, idx, proxy)
                um, _ = ecs.ecs_potential64(X, y, lik, z - dz, idx, proxy)
                np.testing.assert_allclose(g[j], (up - um) / 2e-6, rtol=2e-5, atol=1e-5)
        # with every row in the subsample exactly once the estimate is the full log-likelihood minus the variance correction
        full = np.arange(400, dtype=np.int32)
        proxy = ecs.TaylorProxy.build(X, y, lik, ref, 2)
        z = beta.copy()
        u, _ = ecs.ecs_potential64(X, y, lik, z, full, proxy)
        l, _, _ = ecs._row_loglik(lik, X.astype(np.float64) @ z, y.astype(np.float64))
        l0, d1, d2 = ecs._row_loglik(lik, proxy.eta_ref, y.astype(np.float64))
        a = X.astype(np.float64) @ z - proxy.eta_ref
        diff = l - (l0 + d1 * a + 0.5 * d2 * a * a)
        prior = np.sum(-0.5 * z * z - 0.5 * np.log(2 * np.pi))
        np.testing.assert_allclose(-u, prior + l.sum() - 0.5 * 400 * diff.var(), rtol=1e-10)
        # at the reference point every diff is zero: the estimate is exact whatever the subsample
        u_ref, _ = ecs.ecs_potential64(X, y, lik, ref, idx, proxy)
        np.testing.assert_allclose(-u_ref, np.sum(-0.5 * ref * ref - 0.5 * np.log(2 * np.pi)) + proxy.L0, rtol=1e-12)


def test_randint_and_block_update():
    k = prng.key(3)
    v = ecs.randint(k, 20000, 0, 7)
    assert v.dtype == np.int32 and v.min() == 0 and v.max() == 6
    assert np.abs(np.bincount(v, minlength=7) / 20000 - 1 / 7).max() < 0.01
    assert 0 <= ecs.randint(k, None, 0, 581012) < 581012
    assert ecs.randint(k, None, 5, 6) == 5                                           # span 1
    big = ecs.randint(k, 1000, 0, 2 ** 31 - 1)
    assert big.min() >= 0 and big.max() > 2 ** 29
    # block update (ecs_proxies.py:58-71): exactly one block of ceil(m / num_blocks) positions changes, the rest is kept
    idx = np.arange(100, 123, dtype=np.int32)
    key2, new = ecs.update_block(k, 5, idx, 10 ** 6)
    changed = np.nonzero(new != idx)[0]
    bs = (23 - 1) // 5 + 1
    assert len(changed) <= bs and changed.min() // bs == changed.max() // bs
    assert np.array_equal(key2, prng.split(k, 3)[0])
    # subsample draw (primitives.py:457-469): m distinct rows
    u = ecs.subsample_indices(prng.key(1), 500, 40)
    assert len(set(u.tolist())) == 40 and u.min() >= 0 and u.max() < 500
