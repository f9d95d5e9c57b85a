"""SURVEY.md 8(f) rank 4 on the GPU: ``log_likelihood`` / ``Predictive`` over collected samples (infer/util.py:838-1188)
against the NumPy oracle (oracle/predictive.py), and pickling of ``MCMC`` (mcmc.py:806-809, test_pickle.py:88-95)."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import families, random as b2random                      # noqa: E402
from numpyro_b200.infer import MCMC, NUTS                                    # noqa: E402
from numpyro_b200.predictive import Predictive, log_likelihood               # noqa: E402
from oracle import families as ofam, predictive as opred, prng               # noqa: E402

F = np.float32
J = 8
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])


def _glm_data(n=3000, d=37, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)).astype(F)
    beta = (rng.normal(size=d) * 0.4).astype(F)
    return X, beta, rng


def test_eight_schools_log_likelihood_and_predictive_match_oracle():
    mcmc = MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=100, num_samples=60, num_chains=2,
                chain_method="vectorized", progress_bar=False)
    mcmc.run(b2random.PRNGKey(3), J, S8, y=Y8, extra_fields=("z.tau",))
    samples = mcmc.get_samples()
    fam = ofam.EightSchools(S8, Y8)
    z = np.concatenate([samples["mu"][:, None], np.log(samples["tau"])[:, None], samples["theta_base"]], axis=1).astype(F)
    ll = log_likelihood(families.EightSchoolsNonCentered(), samples, J, S8, y=Y8)["obs"]
    assert ll.shape == (120, 8)
    want = np.stack([opred.log_likelihood(fam, zi) for zi in z])
    np.testing.assert_allclose(ll, want, rtol=2e-5, atol=2e-5)
    # batch_ndims = 2: [chains, samples, ...] (util.py:1146-1147)
    ll2 = log_likelihood(families.EightSchoolsNonCentered(), mcmc.get_samples(group_by_chain=True), J, S8, y=Y8, batch_ndims=2)["obs"]
    np.testing.assert_array_equal(ll2.reshape(120, 8), ll)
    # Predictive: y=None, one key per sample, the observed site gets split(sample_key)[1]
    key = b2random.PRNGKey(11)
    pred = Predictive(families.EightSchoolsNonCentered(), samples)(key, J, S8)
    assert set(pred) == {"obs", "theta"} and pred["obs"].shape == (120, 8)
    skeys = opred.sample_keys(prng.key(11), 120)
    want = np.stack([opred.predictive(fam, zi, k) for zi, k in zip(z, skeys)])
    np.testing.assert_allclose(pred["obs"], want, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(pred["theta"], samples["theta"], rtol=1e-5, atol=1e-5)
    # a single sample uses the key itself (util.py:916-918); return_sites filters
    one = {k: v[:1] for k, v in samples.items()}
    p1 = Predictive(families.EightSchoolsNonCentered(), one, return_sites=["obs"])(key, J, S8)
    assert set(p1) == {"obs"}
    np.testing.assert_allclose(p1["obs"][0], opred.predictive(fam, z[0], prng.key(11)), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_glm_log_likelihood_matches_oracle(lik):
    X, beta, rng = _glm_data()
    eta = X @ beta
    y = (rng.uniform(size=X.shape[0]) < 1 / (1 + np.exp(-eta))).astype(F) if lik == "bernoulli" else rng.poisson(np.exp(np.clip(eta, -4, 3))).astype(F)
    model = families.LogisticRegression() if lik == "bernoulli" else families.PoissonRegression()
    coefs = (beta[None] + rng.normal(size=(21, X.shape[1])) * 0.1).astype(F)
    ll = log_likelihood(model, {"coefs": coefs}, X, y)["obs"]
    fam = ofam.GLM(X, y, likelihood=lik)
    want = np.stack([opred.log_likelihood(fam, c) for c in coefs])
    assert ll.shape == (21, X.shape[0])
    np.testing.assert_allclose(ll, want, rtol=2e-5, atol=2e-5)
    # the sum over observations is the likelihood part of the potential
    np.testing.assert_allclose(ll.sum(axis=1), want.astype(np.float64).sum(axis=1), rtol=1e-5)


def test_logistic_predictive_draws_match_oracle():
    X, beta, rng = _glm_data(n=2500, d=54)
    coefs = (beta[None] + rng.normal(size=(17, 54)) * 0.1).astype(F)
    key = b2random.PRNGKey(7)
    pred = Predictive(families.LogisticRegression(), {"coefs": coefs})(key, X)
    assert pred["obs"].shape == (17, 2500) and pred["obs"].dtype == np.int32
    fam = ofam.GLM(X, np.zeros(2500, F), likelihood="bernoulli")
    skeys = opred.sample_keys(prng.key(7), 17)
    want = np.stack([opred.predictive(fam, c, k) for c, k in zip(coefs, skeys)])
    diff = pred["obs"] != want.astype(np.int32)
    margin = np.stack([opred.bernoulli_margin(fam, c, k) for c, k in zip(coefs, skeys)])
    assert diff.sum() <= 3 and np.all(margin[diff] < 1e-5)           # only ties at fp32 rounding level may differ
    assert 0.3 < pred["obs"].mean() < 0.7


def test_horseshoe_normal_predictive_and_deterministic_sites():
    X, beta, rng = _glm_data(n=800, d=24)
    y = (X @ beta + 0.1 * rng.normal(size=800)).astype(F)
    model = families.HorseshoeRegression("normal")
    S = 9
    post = {"lambdas": np.exp(rng.normal(size=(S, 24)) * 0.3).astype(F), "tau": np.exp(rng.normal(size=(S, 1)) * 0.2).astype(F),
            "unscaled_betas": rng.normal(size=(S, 24)).astype(F), "prec_obs": np.exp(rng.normal(size=S) * 0.2 + 1).astype(F)}
    fam = ofam.horseshoe(X, y, likelihood="normal")
    z = np.zeros((S, fam.dim), F)
    for name, off, size in fam.layout:
        v = post[name].reshape(S, size)
        z[:, off:off + size] = v if name == "unscaled_betas" else np.log(v)
    ll = log_likelihood(model, post, X, y)["Y"]
    np.testing.assert_allclose(ll, np.stack([opred.log_likelihood(fam, zi) for zi in z]), rtol=5e-5, atol=5e-5)
    pred = Predictive(model, post)(b2random.PRNGKey(1), X)
    assert set(pred) == {"Y", "betas"}
    skeys = opred.sample_keys(prng.key(1), S)
    np.testing.assert_allclose(pred["Y"], np.stack([opred.predictive(fam, zi, k) for zi, k in zip(z, skeys)]), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pred["betas"], post["tau"] * post["lambdas"] * post["unscaled_betas"], rtol=1e-5)


def test_errors_follow_the_reference():
    with pytest.raises(ValueError, match="Batch shapes"):            # util.py:1064-1070
        Predictive(families.LogisticRegression(), {"coefs": np.zeros((3, 4), F), "extra": np.zeros((2, 4), F)})
    with pytest.raises(NotImplementedError):
        Predictive(families.LogisticRegression(), None)
    with pytest.raises(ValueError):
        families.LogisticRegression().bind(np.zeros((4, 2), F))     # sampling needs the observations


def test_pickle_mcmc_round_trip():
    """test_pickle.py:88-95: samples survive a pickle round trip; the unpickled object can keep sampling."""
    mcmc = MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=50, num_samples=40, num_chains=2,
                chain_method="vectorized", progress_bar=False)
    mcmc.run(b2random.PRNGKey(0), J, S8, y=Y8)
    clone = pickle.loads(pickle.dumps(mcmc))
    for k, v in mcmc.get_samples().items():
        np.testing.assert_array_equal(clone.get_samples()[k], v)
    assert np.array_equal(clone.last_state.rng_key, mcmc.last_state.rng_key)
    clone.post_warmup_state = clone.last_state
    clone.run(clone.last_state.rng_key, J, S8, y=Y8)                  # continues from the pickled state (mcmc.py:558-587)
    mcmc.post_warmup_state = mcmc.last_state
    mcmc.run(mcmc.last_state.rng_key, J, S8, y=Y8)
    np.testing.assert_array_equal(clone.get_samples()["mu"], mcmc.get_samples()["mu"])
    kernel = pickle.loads(pickle.dumps(NUTS(families.LogisticRegression(), max_tree_depth=(5, 7))))
    assert kernel._cfg["max_tree_depth"] == 7 and kernel._engine is None
