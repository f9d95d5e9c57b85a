"""Dense / user-supplied mass matrices on the GPU (numpyro/infer/hmc_util.py:192-237, 439-515,
726-728, 1193-1194; test/infer/test_mcmc.py:75-101, 312-345): whole runs bit-exact against the oracle in the warp and the
streaming regime, the HMCAdaptState matrices, continuing from a dense state, and the public ``dense_mass=True`` API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng, families as model_families, random as b2random      # noqa: E402
from numpyro_b200.infer import MCMC, NUTS                                                         # noqa: E402
from oracle import chain, families, prng                                                          # noqa: E402
from test_gpu_parity import FIELDS, S8, Y8, assert_run_equal, device_potential, eight_schools     # noqa: E402

F = np.float32


def _spd(d, seed):
    a = np.random.default_rng(seed).normal(size=(d, d))
    return (a @ a.T / d + 0.5 * np.eye(d)).astype(F)


def test_dense_mass_eight_schools_bit_exact_and_adapt_state():
    e = eight_schools(3, dense_mass=1)
    keys = prng.split(prng.key(3), 3)
    e.init(keys, 150)
    out = e.run(190, 150, fields=FIELDS)
    fam = families.EightSchools(S8, Y8)
    last = None
    for c in (0, 2):
        kern = chain.Kernel(device_potential(e, c), dense_mass=True)
        res, last = chain.run_chain(kern, fam, keys[c], 150, 40, fields=FIELDS)
        assert_run_equal(out, res, c)
    ds = e.dense_state()
    a = last.adapt_state
    np.testing.assert_array_equal(ds["inverse_mass_matrix"][2], a.inverse_mass_matrix)
    np.testing.assert_array_equal(ds["mass_matrix_sqrt"][2], a.mass_matrix_sqrt)
    np.testing.assert_array_equal(ds["mass_matrix_sqrt_inv"][2], a.mass_matrix_sqrt_inv)
    assert np.abs(a.inverse_mass_matrix - np.diag(np.diag(a.inverse_mass_matrix))).max() > 1e-3      # genuinely dense
    e.close()


@pytest.mark.parametrize("dense", [False, True])
def test_user_supplied_inverse_mass_matrix_bit_exact(dense):
    e = eight_schools(2, dense_mass=int(dense), adapt_mass_matrix=0)
    imm = _spd(10, 5) if dense else np.exp(np.random.default_rng(5).normal(size=10) * 0.3).astype(F)
    e.set_inverse_mass_matrix(imm)
    keys = prng.split(prng.key(11), 2)
    e.init(keys, 60)
    out = e.run(80, 60, fields=FIELDS)
    fam = families.EightSchools(S8, Y8)
    kern = chain.Kernel(device_potential(e, 1), dense_mass=dense, adapt_mass_matrix=False)
    res, _ = chain.run_chain(kern, fam, keys[1], 60, 20, fields=FIELDS, inverse_mass_matrix=imm)
    assert_run_equal(out, res, 1)
    e.close()


def test_dense_mass_tall_data_gemm_regime_bit_exact():
    """Tall-data logistic regression: a dense kernel runs in the GEMM regime (the streaming regime compiles the dense branches
    out of its tick and refuses dense handles)."""
    rng = np.random.default_rng(2)
    N, D, C = 60000, 20, 4
    X = rng.normal(size=(N, D)).astype(F)
    X[:, 1] = (0.8 * X[:, 0] + 0.6 * X[:, 1]).astype(F)                  # correlated columns -> correlated posterior
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ (rng.normal(size=D) * 0.3)))).astype(F)
    with pytest.raises(eng.EngineError, match="streaming regime"):
        eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, regime=_capi.REGIME_STREAM, dense_mass=1)
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, dense_mass=1, max_tree_depth_warmup=6, max_tree_depth=6)
    assert e.regime == _capi.REGIME_GEMM                                   # what "auto" picks for tall data + dense_mass
    keys = prng.split(prng.key(4), C)
    e.init(keys, 120)
    out = e.run(140, 120, fields=FIELDS)
    fam = families.logistic_regression(X, y)
    kern = chain.Kernel(device_potential(e, 1), dense_mass=True, max_tree_depth=(6, 6))
    res, last = chain.run_chain(kern, fam, keys[1], 120, 20, fields=FIELDS)
    assert_run_equal(out, res, 1)
    imm = e.dense_state()["inverse_mass_matrix"][1]
    np.testing.assert_array_equal(imm, last.adapt_state.inverse_mass_matrix)
    cor = imm[0, 1] / np.sqrt(imm[0, 0] * imm[1, 1])
    assert cor < -0.3                                                     # the adapted matrix sees the posterior correlation
    e.close()


def test_dense_mass_public_api_and_resume():
    """MCMC(NUTS(model, dense_mass=True)): adapt-state layout of hmc.py:759-769 (one block over all sites), continuing from
    last_state (mcmc.py:558-587) and the same posterior as the diagonal kernel."""
    J = 8
    kw = dict(num_warmup=300, num_samples=400, num_chains=4, chain_method="vectorized", progress_bar=False)
    m = MCMC(NUTS(model_families.EightSchoolsNonCentered(), dense_mass=True), **kw)
    m.run(b2random.PRNGKey(0), J, S8, y=Y8)
    s = m.get_samples()
    st = m.last_state
    key = ("mu", "tau", "theta_base")
    assert st.adapt_state.inverse_mass_matrix[key].shape == (4, 10, 10)
    assert st.adapt_state.mass_matrix_sqrt[key].shape == (4, 10, 10) and st.adapt_state.mm_state[key][1].shape == (4, 10, 10)
    sq, sqi = st.adapt_state.mass_matrix_sqrt[key][0].astype(np.float64), st.adapt_state.mass_matrix_sqrt_inv[key][0].astype(np.float64)
    np.testing.assert_allclose(sq @ sqi, np.eye(10), atol=1e-3)
    np.testing.assert_allclose(sqi.T @ sqi, st.adapt_state.inverse_mass_matrix[key][0], rtol=1e-3, atol=1e-4)
    d = MCMC(NUTS(model_families.EightSchoolsNonCentered()), **kw)
    d.run(b2random.PRNGKey(0), J, S8, y=Y8)
    sd = d.get_samples()
    assert abs(s["mu"].mean() - sd["mu"].mean()) < 0.6 and abs(s["tau"].mean() - sd["tau"].mean()) < 0.6
    # resume: a longer run == run + continue from last_state
    long = MCMC(NUTS(model_families.EightSchoolsNonCentered(), dense_mass=True), num_warmup=300, num_samples=450, num_chains=4,
                chain_method="vectorized", progress_bar=False)
    long.run(b2random.PRNGKey(0), J, S8, y=Y8)
    m.num_samples = 50
    m.post_warmup_state = m.last_state
    m.run(m.last_state.rng_key, J, S8, y=Y8)
    np.testing.assert_array_equal(m.get_samples(group_by_chain=True)["mu"], long.get_samples(group_by_chain=True)["mu"][:, 400:])
    with pytest.raises(NotImplementedError):
        NUTS(model_families.EightSchoolsNonCentered(), dense_mass=[("mu", "tau")])
    # supplied matrix through the kernel argument
    u = MCMC(NUTS(model_families.EightSchoolsNonCentered(), dense_mass=True, inverse_mass_matrix=_spd(10, 1), adapt_mass_matrix=False),
             num_warmup=50, num_samples=20, num_chains=2, chain_method="vectorized", progress_bar=False)
    u.run(b2random.PRNGKey(1), J, S8, y=Y8)
    np.testing.assert_array_equal(u.last_state.adapt_state.inverse_mass_matrix[key][1], _spd(10, 1))
