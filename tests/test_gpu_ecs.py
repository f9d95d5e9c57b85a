"""HMCECS on the GPU (numpyro/infer/hmc_gibbs.py:502-690, contrib/ecs_proxies.py, the default algorithm of
examples/covtype.py:154-165): the subsampled potential against the fp64 oracle, a whole HMCECS chain (block updates, Metropolis
tests of the subsample, inner NUTS transitions with adaptation) bit-exact against the oracle driven by the engine's own
potential, and the public ``MCMC(HMCECS(NUTS(model), ...))`` call against full-data NUTS."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng, families as model_families, random as b2random      # noqa: E402
from numpyro_b200.hmc_gibbs import HMCECS, _randint, _subsample_indices, _update_block              # noqa: E402
from numpyro_b200.infer import MCMC, NUTS                                                         # noqa: E402
from oracle import ecs, families, prng                                                            # noqa: E402

F = np.float32


def _data(n, d, lik="bernoulli", seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.normal(size=(n, d)) * (0.3 if lik == "poisson" else 1.0)).astype(F)
    beta = rng.normal(size=d) * 0.4
    eta = X.astype(np.float64) @ beta
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-eta))).astype(F) if lik == "bernoulli" else rng.poisson(np.exp(eta)).astype(F)
    return X, y, beta


def _map_estimate(X, y, iters=12):
    X64, y64 = X.astype(np.float64), y.astype(np.float64)
    b = np.zeros(X.shape[1])
    for _ in range(iters):
        p = 1 / (1 + np.exp(-(X64 @ b)))
        H = (X64 * (p * (1 - p))[:, None]).T @ X64 + np.eye(X.shape[1])
        b = b - np.linalg.solve(H, X64.T @ (p - y64) + b)
    return b.astype(F)


def _engine(X, y, C, m, degree, lik="bernoulli", **kw):
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, ecs_subsample_size=m, ecs_proxy_degree=degree,
                   likelihood=_capi.LIK_POISSON_LOG if lik == "poisson" else _capi.LIK_BERNOULLI_LOGIT, **kw)
    assert e.regime == _capi.REGIME_WARP
    return e


def _install(e, X, y, lik, ref, degree):
    p = ecs.TaylorProxy.build(X, y, lik, ref, degree)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, F)).to(e.device)
    e.ecs_set_proxy(dev(p.ref), dev(p.eta_ref), dev(p.G), dev(p.H), float(p.L0))
    return p


def test_integer_draws_match_the_oracle():
    keys = prng.split(prng.key(5), 6)
    np.testing.assert_array_equal(_randint(keys, 9, 0, 581012), np.stack([ecs.randint(k, 9, 0, 581012) for k in keys]))
    np.testing.assert_array_equal(_randint(keys, None, 0, np.arange(6) + 3), np.array([ecs.randint(k, None, 0, 3 + i) for i, k in enumerate(keys)]))
    np.testing.assert_array_equal(_subsample_indices(keys[0], 5000, 37), ecs.subsample_indices(keys[0], 5000, 37))
    idx = np.stack([np.arange(50, dtype=np.int32) + 100 * c for c in range(6)])
    got = _update_block(keys, 7, idx, 10 ** 5)
    for c in range(6):
        np.testing.assert_array_equal(got[c], ecs.update_block(keys[c], 7, idx[c], 10 ** 5)[1])


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
@pytest.mark.parametrize("degree", [0, 1, 2])
def test_subsampled_potential_matches_fp64_oracle(lik, degree):
    N, D, C, m = 20000, 23, 5, 300
    X, y, beta = _data(N, D, lik)
    ref = (beta + np.random.default_rng(1).normal(size=D) * 0.02).astype(F)
    e = _engine(X, y, C, m, degree, lik)
    p = _install(e, X, y, lik, ref, degree) if degree else None
    rng = np.random.default_rng(2)
    idx = rng.integers(0, N, size=(C, m)).astype(np.int32)
    e.ecs_set_indices(idx)
    z = (ref[None] + rng.normal(size=(C, D)) * 0.03).astype(F)
    U, g = e.potential_and_grad(z)
    U, g = U.cpu().numpy(), g.cpu().numpy()
    for c in range(C):
        u64, g64 = ecs.ecs_potential64(X, y, lik, z[c].astype(np.float64), idx[c], p)
        np.testing.assert_allclose(U[c], u64, rtol=2e-5)
        np.testing.assert_allclose(g[c], g64, rtol=2e-3, atol=2e-4 * np.abs(g64).max())
    e.close()


def test_hmcecs_chain_bit_exact_against_oracle():
    """Engine-level replay of HMCECS.init / sample (hmc_gibbs.py:577-682) for two chains; the oracle's inner kernel asks the
    engine's own potential hook (a second handle), so every integer and every bit of the bookkeeping must agree."""
    N, D, C, m, W, S, NB = 6000, 9, 2, 120, 25, 15, 6
    X, y, beta = _data(N, D, seed=3)
    ref = _map_estimate(X, y)
    kernel = HMCECS(NUTS(model_families.LogisticRegression(), max_tree_depth=5), num_blocks=NB, proxy=HMCECS.taylor_proxy({"coefs": ref}))
    keys = prng.split(prng.key(4), C)
    state = kernel.init(keys, W, None, (X, y, m), {})
    hook = _engine(X, y, C, m, 2)
    _install(hook, X, y, "bernoulli", ref, 2)

    def potential_at_chain(c):
        def potential_at(u):
            def pot(z):
                idx = np.zeros((C, m), np.int32); idx[c] = u
                hook.ecs_set_indices(idx)
                zz = np.zeros((C, D), F); zz[c] = z
                U, g = hook.potential_and_grad(zz)
                return F(U[c].item()), g[c].cpu().numpy()
            return pot
        return potential_at
    fam = families.logistic_regression(X, y)
    oracles, ostates = [], []
    for c in range(C):
        o = ecs.HMCECS(dict(max_tree_depth=(5, 5)), potential_at_chain(c), N, m, NB)
        oracles.append(o)
        ostates.append(o.init(keys[c], W, fam))
        np.testing.assert_array_equal(state.z["N"][c], ostates[c].u)
        np.testing.assert_array_equal(state.z["coefs"][c], ostates[c].hmc_state.z)
    accepted = 0
    for i in range(W + S):
        state = kernel.sample(state, (X, y, m), {})
        for c in range(C):
            prev_u = ostates[c].u
            ostates[c] = oracles[c].sample(ostates[c])
            accepted += int(not np.array_equal(prev_u, ostates[c].u))
            np.testing.assert_array_equal(state.z["N"][c], ostates[c].u, err_msg=f"step {i} chain {c}")
            np.testing.assert_array_equal(state.z["coefs"][c], ostates[c].hmc_state.z, err_msg=f"step {i} chain {c}")
            assert state.accept_prob[c] == ostates[c].accept_prob
            assert state.hmc_state.num_steps[c] == ostates[c].hmc_state.num_steps
            assert state.hmc_state.adapt_state.step_size[c] == ostates[c].hmc_state.adapt_state.step_size
            np.testing.assert_array_equal(state.rng_key[c], ostates[c].rng_key)
    assert 0 < accepted < 2 * (W + S)                       # both outcomes of the subsample's Metropolis test occurred
    hook.close()


def test_hmcecs_public_api_posterior_close_to_full_data_nuts():
    """examples/covtype.py:154-165 at a small shape: MCMC(HMCECS(NUTS(model), num_blocks, proxy=taylor_proxy(MAP)))."""
    N, D, m = 20000, 6, 400
    X, y, beta = _data(N, D, seed=7)
    ref = _map_estimate(X, y)
    full = MCMC(NUTS(model_families.LogisticRegression()), num_warmup=300, num_samples=400, num_chains=2, chain_method="vectorized")
    full.run(b2random.PRNGKey(0), X, y)
    want = full.get_samples()["coefs"]
    kern = HMCECS(NUTS(model_families.LogisticRegression()), num_blocks=10, proxy=HMCECS.taylor_proxy({"coefs": ref}))
    mc = MCMC(kern, num_warmup=300, num_samples=400, num_chains=2, chain_method="vectorized")
    mc.run(b2random.PRNGKey(1), X, y, m, extra_fields=("accept_prob", "hmc_state.accept_prob"))
    got = mc.get_samples()["coefs"]
    assert got.shape == (800, D) and set(mc.get_samples()) == {"coefs"}
    sd = want.std(0)
    assert np.all(np.abs(got.mean(0) - want.mean(0)) < 0.5 * sd), (got.mean(0), want.mean(0), sd)
    assert np.all(np.abs(got.std(0) / sd - 1) < 0.35)
    acc = mc.get_extra_fields()["accept_prob"]
    assert acc.shape == (800,) and 0.2 < acc.mean() <= 1.0
    assert mc.get_extra_fields()["hmc_state.accept_prob"].mean() > 0.5
    with pytest.raises(AssertionError):                      # no subsample statement (hmc_gibbs.py:596)
        MCMC(HMCECS(NUTS(model_families.LogisticRegression())), num_warmup=1, num_samples=1).run(b2random.PRNGKey(0), X, y)
