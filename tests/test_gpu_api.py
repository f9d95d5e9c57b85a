"""numpyro-style front end (numpyro_b200.infer.MCMC / NUTS) on the GPU: collection layout, extra
fields, warm-up/run split, post_warmup_state -- the behaviours pinned by the reference's
test/infer/test_mcmc.py (:437-485, :512-528, :688-700, :1231-1251)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import families, random as b2random            # noqa: E402
from numpyro_b200.infer import MCMC, NUTS, HMC                     # noqa: E402

J = 8
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])


def test_eight_schools_api_shapes_and_extra_fields():
    mcmc = MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=200, num_samples=300, num_chains=4,
                chain_method="vectorized", progress_bar=False)
    mcmc.run(b2random.PRNGKey(0), J, S8, y=Y8,
             extra_fields=("potential_energy", "num_steps", "adapt_state.step_size", "z.tau", "~z.theta_base"))
    s = mcmc.get_samples(group_by_chain=True)
    assert set(s) == {"mu", "tau", "theta"}                        # test_mcmc.py:512-528 + '~z.site' removal
    assert s["mu"].shape == (4, 300) and s["theta"].shape == (4, 300, 8)
    flat = mcmc.get_samples()
    assert flat["theta"].shape == (1200, 8)
    ex = mcmc.get_extra_fields(group_by_chain=True)
    assert ex["potential_energy"].shape == (4, 300) and ex["diverging"].dtype == bool
    np.testing.assert_allclose(s["tau"], np.exp(ex["z.tau"]), rtol=1e-6)     # test_mcmc.py:1245-1251
    assert np.all(ex["adapt_state.step_size"][:, 0] == ex["adapt_state.step_size"][:, -1])   # frozen after warm-up
    st = mcmc.last_state
    assert st.i.shape == (4,) and np.all(st.i == 500) and st.z["theta_base"].shape == (4, 8)
    assert st.adapt_state.inverse_mass_matrix[("mu", "tau", "theta_base")].shape == (4, 10)
    mcmc.print_summary()
    assert abs(flat["mu"].mean() - 4.4) < 1.0 and abs(flat["tau"].mean() - 3.6) < 1.0


def test_warmup_then_run_matches_single_run_and_state_resume():
    """test_mcmc.py:437-485."""
    key = b2random.PRNGKey(2)
    kw = dict(num_warmup=100, num_samples=50, num_chains=2, chain_method="vectorized", progress_bar=False)
    a = MCMC(NUTS(families.EightSchoolsNonCentered()), **kw)
    a.warmup(key, J, S8, y=Y8)
    wstate = a.post_warmup_state
    assert np.all(wstate.i == 100)
    a.run(wstate.rng_key, J, S8, y=Y8)
    first = a.get_samples()
    b = MCMC(NUTS(families.EightSchoolsNonCentered()), **kw)
    b.run(key, J, S8, y=Y8)
    np.testing.assert_array_equal(first["mu"], b.get_samples()["mu"])
    # continue sampling from the last state (mcmc.py:558-587): num_samples MORE draws, bit-identical to the tail of one
    # longer run because every chain keeps its own key stream
    long = MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=100, num_samples=100, num_chains=2,
                chain_method="vectorized", progress_bar=False)
    long.run(key, J, S8, y=Y8)
    want = long.get_samples(group_by_chain=True)
    b.post_warmup_state = b.last_state
    b.run(b.last_state.rng_key, J, S8, y=Y8)
    got = b.get_samples(group_by_chain=True)
    assert np.all(b.last_state.i == 200)
    for k in want:
        assert got[k].shape == (2, 50) + want[k].shape[2:]
        np.testing.assert_array_equal(got[k], want[k][:, 50:])
        assert np.all(np.isfinite(got[k])) and np.std(got[k]) > 0
    assert not np.array_equal(got["tau"], first["tau"].reshape(2, 50))
    # the reference pattern with a FRESH MCMC object (no engine yet when the state is assigned)
    c = MCMC(NUTS(families.EightSchoolsNonCentered()), **kw)
    c.post_warmup_state = a.last_state
    c.run(a.last_state.rng_key, J, S8, y=Y8)
    np.testing.assert_array_equal(c.get_samples(group_by_chain=True)["tau"], want["tau"][:, 50:])
    # a second run() after warmup() + run() restarts from the post warm-up state (mcmc.py:677-679)
    a.run(wstate.rng_key, J, S8, y=Y8)
    np.testing.assert_array_equal(a.get_samples()["mu"], first["mu"])


def test_mcmc_kernel_interface_init_sample_postprocess():
    """The MCMCKernel plug-in surface (mcmc.py:79-124): kernel.init + repeated kernel.sample reproduce MCMC.run
    transition by transition; postprocess_fn constrains like the collection path."""
    key = b2random.PRNGKey(3)
    nw, ns = 40, 15
    ref = MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=nw, num_samples=ns, num_chains=1, progress_bar=False)
    ref.run(key, J, S8, y=Y8, extra_fields=("num_steps", "z.tau"))
    want = ref.get_samples()
    kern = NUTS(families.EightSchoolsNonCentered())
    assert kern.sample_field == "z" and kern.default_fields == ("z", "diverging") and not kern.is_ensemble_kernel
    state = kern.init(key, nw, None, model_args=(J, S8), model_kwargs=dict(y=Y8))
    assert int(state.i) == 0 and state.z["theta_base"].shape == (8,) and state.adapt_state.step_size == 1.0
    post = kern.postprocess_fn((J, S8), dict(y=Y8))
    steps = []
    for t in range(nw + ns):
        state = kern.sample(state, (J, S8), dict(y=Y8))
        assert int(state.i) == t + 1
        if t >= nw:
            con = post(state.z)
            np.testing.assert_array_equal(con["mu"], want["mu"][t - nw])
            np.testing.assert_array_equal(con["theta"], want["theta"][t - nw])
            np.testing.assert_allclose(con["tau"], np.exp(state.z["tau"]), rtol=1e-6)
            steps.append(int(state.num_steps))
    np.testing.assert_array_equal(steps, ref.get_extra_fields()["num_steps"])
    assert "steps of size" in kern.get_diagnostics_str(state)
    # a state the engine does not hold (edited by an outer kernel such as HMCGibbs) is loaded before the transition
    edited = ref.last_state
    again = kern.sample(edited, (J, S8), dict(y=Y8))
    assert int(again.i) == nw + ns + 1
    # vectorised: a batch of keys gives a batched state
    vstate = kern.init(b2random.split(key, 3), 5, None, model_args=(J, S8), model_kwargs=dict(y=Y8))
    assert vstate.z["mu"].shape == (3,) and kern.sample(vstate, (J, S8), dict(y=Y8)).i.tolist() == [1, 1, 1]


def test_init_strategies():
    """initialization.py:88-155."""
    from numpyro_b200.infer import init_to_feasible, init_to_uniform, init_to_value
    key = b2random.PRNGKey(4)
    kw = dict(num_warmup=0, num_samples=1, num_chains=2, chain_method="vectorized", progress_bar=False)
    vals = {"mu": 1.5, "tau": 2.0, "theta_base": np.linspace(-1, 1, 8)}
    z0 = {"mu": np.full(2, 1.5), "tau": np.full((2,), np.log(2.0)), "theta_base": np.tile(np.linspace(-1, 1, 8), (2, 1))}
    a = MCMC(NUTS(families.EightSchoolsNonCentered(), init_strategy=init_to_value(values=vals)), **kw)
    a.run(key, J, S8, y=Y8)
    b = MCMC(NUTS(families.EightSchoolsNonCentered()), **kw)
    b.run(key, J, S8, y=Y8, init_params=z0)
    np.testing.assert_array_equal(a.get_samples()["theta"], b.get_samples()["theta"])      # same start, same key stream
    f = NUTS(families.EightSchoolsNonCentered(), init_strategy=init_to_feasible())
    st = f.init(key, 0, None, model_args=(J, S8), model_kwargs=dict(y=Y8))
    assert np.all(st.z["theta_base"] == 0) and st.z["mu"] == 0 and st.z["tau"] == 0
    u = NUTS(families.EightSchoolsNonCentered(), init_strategy=init_to_uniform(radius=0.5))
    st = u.init(key, 0, None, model_args=(J, S8), model_kwargs=dict(y=Y8))
    assert np.all(np.abs(st.z["theta_base"]) <= 0.5) and np.any(st.z["theta_base"] != 0)
    with pytest.raises(NotImplementedError):
        MCMC(NUTS(families.EightSchoolsNonCentered(), init_strategy=init_to_value(values={"mu": 0.0})), **kw).run(key, J, S8, y=Y8)


def test_hmc_and_thinning_and_sequential_chains():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(500, 3)).astype(np.float32)
    y = (rng.uniform(size=500) < 1 / (1 + np.exp(-(X @ np.array([1.0, -1.0, 0.5]))))).astype(np.float32)
    vec = MCMC(HMC(families.LogisticRegression(), num_steps=8, step_size=0.05), num_warmup=100, num_samples=90,
               num_chains=2, thinning=3, chain_method="vectorized", progress_bar=False)
    vec.run(b2random.PRNGKey(1), X, y)
    seq = MCMC(HMC(families.LogisticRegression(), num_steps=8, step_size=0.05), num_warmup=100, num_samples=90,
               num_chains=2, thinning=3, chain_method="sequential", progress_bar=False)
    seq.run(b2random.PRNGKey(1), X, y)
    a, b = vec.get_samples(group_by_chain=True)["coefs"], seq.get_samples(group_by_chain=True)["coefs"]
    assert a.shape == (2, 30, 3)
    np.testing.assert_array_equal(a, b)                            # chains are independent of how they are batched
    assert np.all(np.abs(a.mean(axis=(0, 1)) - np.array([1.0, -1.0, 0.5])) < 0.5)


def test_horseshoe_api_deterministic_site():
    rng = np.random.default_rng(1)
    X = rng.normal(size=(100, 8)).astype(np.float32)
    Y = (X[:, 0] * 2 - X[:, 1] + 0.5 * X[:, 2] + 0.05 * rng.normal(size=100)).astype(np.float32)
    mcmc = MCMC(NUTS(families.HorseshoeRegression("normal")), num_warmup=300, num_samples=300, num_chains=2,
                chain_method="vectorized", progress_bar=False)
    mcmc.run(b2random.PRNGKey(0), X, Y)
    s = mcmc.get_samples()
    assert set(s) == {"lambdas", "tau", "unscaled_betas", "prec_obs", "betas"}
    np.testing.assert_allclose(s["betas"], s["tau"] * s["lambdas"] * s["unscaled_betas"], rtol=2e-5, atol=1e-6)
    b = s["betas"].mean(axis=0)
    assert abs(b[0] - 2.0) < 0.1 and abs(b[1] + 1.0) < 0.1 and abs(b[2] - 0.5) < 0.1 and np.all(np.abs(b[3:]) < 0.1)
