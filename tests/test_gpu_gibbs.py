"""HMCGibbs on the GPU (numpyro/infer/hmc_gibbs.py:38-192): eight schools with ``mu`` resampled from its
exact conditional by a user ``gibbs_fn`` and NUTS (the engine, conditioned on ``mu``) for tau / theta_base -- the conditioned
potential against the full one, a whole chain bit-exact against the oracle, and the posterior against plain NUTS."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng, families as model_families, random as b2random      # noqa: E402
from numpyro_b200.hmc_gibbs import HMCGibbs                                                       # noqa: E402
from numpyro_b200.infer import MCMC, NUTS                                                         # noqa: E402
from oracle import ecs, families, prng                                                            # noqa: E402
from test_gpu_parity import S8, Y8                                                                 # noqa: E402

F = np.float32
J = 8


def gibbs_mu(rng_key, gibbs_sites, hmc_sites):
    """mu | tau, theta_base, y ~ Normal(M, P^-1/2): prior N(0, 5) times the eight Normal likelihood terms."""
    tau, tb = np.float64(np.ravel(hmc_sites["tau"])[0]), np.asarray(hmc_sites["theta_base"], np.float64)
    prec = 1.0 / 25.0 + np.sum(1.0 / S8.astype(np.float64) ** 2)
    mean = np.sum((Y8 - tau * tb) / S8.astype(np.float64) ** 2) / prec
    return {"mu": F(mean + eng.prng_normal(rng_key, 1)[0] / np.sqrt(prec))}


class _Reduced:
    """Site table of the inner kernel (trace order / flat order) when mu is a Gibbs site."""
    init_sites = [("tau", 1), ("theta_base", J)]
    layout = [("tau", 0, 1), ("theta_base", 1, J)]


def _conditioned_engine(C):
    fixed = np.zeros(J + 2, np.int32)
    fixed[0] = 1
    return eng.Engine(family=_capi.FAMILY_EIGHT_SCHOOLS, num_chains=C, n_rows=J, y=Y8, aux=S8, tau_scale=5.0, cond_fixed=fixed)


def test_conditioned_potential_is_the_full_potential_with_mu_substituted():
    C = 3
    e = _conditioned_engine(C)
    assert e.D == J + 1 and e.Dfull == J + 2 and e.regime == _capi.REGIME_WARP
    rng = np.random.default_rng(0)
    mu = rng.normal(size=C).astype(F) * 3
    vals = np.zeros((C, J + 3), F)
    vals[:, 0] = mu
    e.cond_set_values(vals)
    z = (rng.normal(size=(C, J + 1)) * 0.5).astype(F)
    U, g = e.potential_and_grad(z)
    fam = families.EightSchools(S8, Y8)
    for c in range(C):
        u64, g64 = fam.potential64(np.concatenate([[mu[c]], z[c]]).astype(np.float64))
        np.testing.assert_allclose(U[c].item(), u64, rtol=1e-5)
        np.testing.assert_allclose(g[c].cpu().numpy(), g64[1:], rtol=1e-5, atol=1e-5)
    e.close()


def test_hmcgibbs_chain_bit_exact_against_oracle():
    C, W, S = 2, 40, 20
    kernel = HMCGibbs(NUTS(model_families.EightSchoolsNonCentered(), max_tree_depth=6), gibbs_fn=gibbs_mu, gibbs_sites=["mu"])
    keys = prng.split(prng.key(6), C)
    state = kernel.init(keys, W, None, (J, S8), {"y": Y8})
    hook = _conditioned_engine(C)

    def potential_at_chain(c):
        def potential_at(gibbs):
            def pot(z):
                vals = np.zeros((C, J + 3), F); vals[c, 0] = gibbs["mu"]
                hook.cond_set_values(vals)
                zz = np.zeros((C, J + 1), F); zz[c] = z
                U, g = hook.potential_and_grad(zz)
                return F(U[c].item()), g[c].cpu().numpy()
            return pot
        return potential_at

    def constrain_hmc(z, gibbs):
        # (the engine's own transform: exp of the constraint is not part of the det-f32 convention)
        full = np.concatenate([[gibbs["mu"]], z]).astype(F)[None]
        con = hook.constrain(torch.from_numpy(full).to(hook.device)).cpu().numpy()[0]
        return {"tau": con[1:2], "theta_base": con[2:2 + J]}

    def prior_draw(key_u):
        k_mu = prng.split(key_u)[1]                                     # first latent site of the trace
        return {"mu": F(F(5.0) * prng.normal(k_mu))}
    oracles = [ecs.HMCGibbs(dict(max_tree_depth=(6, 6)), potential_at_chain(c), gibbs_mu, constrain_hmc, prior_draw) for c in range(C)]
    ostates = [o.init(keys[c], W, _Reduced()) for c, o in enumerate(oracles)]
    for c in range(C):
        assert state.z["mu"][c] == ostates[c].gibbs["mu"]
        np.testing.assert_array_equal(np.concatenate([np.atleast_1d(state.z["tau"][c]), state.z["theta_base"][c]]), ostates[c].hmc_state.z)
    for i in range(W + S):
        state = kernel.sample(state, (J, S8), {"y": Y8})
        for c in range(C):
            ostates[c] = oracles[c].sample(ostates[c])
            assert state.z["mu"][c] == ostates[c].gibbs["mu"], (i, c)
            np.testing.assert_array_equal(np.concatenate([np.atleast_1d(state.z["tau"][c]), state.z["theta_base"][c]]), ostates[c].hmc_state.z, err_msg=f"step {i} chain {c}")
            assert state.hmc_state.num_steps[c] == ostates[c].hmc_state.num_steps
            np.testing.assert_array_equal(state.rng_key[c], ostates[c].rng_key)
    hook.close()


def test_hmcgibbs_public_api_posterior_matches_plain_nuts():
    kw = dict(num_warmup=400, num_samples=600, num_chains=4, chain_method="vectorized")
    plain = MCMC(NUTS(model_families.EightSchoolsNonCentered()), **kw)
    plain.run(b2random.PRNGKey(0), J, S8, y=Y8)
    want = plain.get_samples()
    mc = MCMC(HMCGibbs(NUTS(model_families.EightSchoolsNonCentered()), gibbs_fn=gibbs_mu, gibbs_sites=["mu"]), **kw)
    mc.run(b2random.PRNGKey(1), J, S8, y=Y8)
    got = mc.get_samples()
    assert set(got) == {"mu", "tau", "theta_base", "theta"} and got["theta"].shape == (2400, J)
    for name in ("mu", "tau"):
        assert abs(got[name].mean() - want[name].mean()) < 0.6, (name, got[name].mean(), want[name].mean())
    np.testing.assert_allclose(got["theta"], got["mu"][:, None] + got["tau"][:, None] * got["theta_base"], rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError):
        HMCGibbs(NUTS(model_families.EightSchoolsNonCentered()), gibbs_fn=gibbs_mu, gibbs_sites=["nope"]).init(b2random.PRNGKey(0), 5, None, (J, S8), {"y": Y8})
