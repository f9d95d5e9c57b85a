"""Parity at the SHAPES of BASELINE configs 3 and 4 (chains / rows reduced so the oracle finishes in seconds):
  config 3: hierarchical GLM, D = 256 columns (a block of group columns shares a global scale), many chains,
            Bernoulli-logit and Poisson-log likelihoods;
  config 4: horseshoe regression, D = 1000 columns (latent 2001 / 2002), Normal and Bernoulli likelihoods.
Today these run in the warp-per-chain regime (the tcgen05 many-chain regime is the next build step);
the tests pin what that regime must reproduce: potentials rtol 1e-5 against fp64 and whole runs bit-exact against
the oracle driven by the engine's own potential hook."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng            # noqa: E402
from oracle import chain, families, prng                 # noqa: E402
from test_gpu_parity import FIELDS, assert_run_equal, device_potential, glm_engine      # noqa: E402

F = np.float32


def _check(e, fam, z, chains):
    U, g = e.potential_and_grad(z)
    U, g = U.cpu().numpy(), g.cpu().numpy()
    for c in chains:
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U[c], u64, rtol=1e-5)
        np.testing.assert_allclose(g[c], g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max())


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_config3_shape_hierarchical_glm_many_chains(lik):
    rng = np.random.default_rng(33)
    N, D, C = 1500, 256, 96
    X = (rng.normal(size=(N, D)) / np.sqrt(D)).astype(F)
    X[:, 192:] = 0.0
    X[np.arange(N), 192 + rng.integers(0, 64, size=N)] = 1.0          # one-hot group block: random intercepts
    beta = rng.normal(size=D) * 0.5
    eta = np.clip(X @ beta, -10, 10)
    if lik == "bernoulli":
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(F)
        kw, okw = {}, {}
    else:
        y = rng.poisson(np.exp(eta)).astype(F)
        kw, okw = dict(likelihood=_capi.LIK_POISSON_LOG), dict(likelihood="poisson")
    e = glm_engine(C, X, y, global_scale=_capi.SCALE_HALFCAUCHY, group_col_begin=192, group_col_end=256, tau_scale=1.0,
                   max_tree_depth_warmup=4, max_tree_depth=4, **kw)
    assert e.regime == _capi.REGIME_WARP and e.D == D + 1
    fam = families.GLM(X, y, global_scale="halfcauchy", group_cols=(192, 256), tau_scale=1.0, **okw)
    z = (rng.normal(size=(C, e.D)) * 0.3).astype(F)
    _check(e, fam, z, (0, 41, C - 1))
    keys = prng.split(prng.key(5), C)
    e.init(keys, 12)
    out = e.run(20, 12, fields=FIELDS)
    for c in (7, C - 1):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(4, 4))
        res, _ = chain.run_chain(kern, fam, keys[c], 12, 8, fields=FIELDS)
        assert_run_equal(out, res, c)
    assert int(out["num_steps"].sum().item()) >= C * 8


@pytest.mark.parametrize("lik", ["normal", "bernoulli"])
def test_config4_shape_horseshoe_thousand_columns(lik):
    rng = np.random.default_rng(44)
    N, D, C = 300, 1000, 6
    X = rng.normal(size=(N, D)).astype(F)
    X -= X.mean(0)
    eta = 2 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2]                       # examples/horseshoe_regression.py:105-125
    if lik == "normal":
        y = (eta + 0.05 * rng.normal(size=N)).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                       max_tree_depth_warmup=4, max_tree_depth=4)
        assert e.D == 2 * D + 2
    else:
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(F)
        e = glm_engine(C, X, y, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY, max_tree_depth_warmup=4, max_tree_depth=4)
        assert e.D == 2 * D + 1
    fam = families.horseshoe(X, y, lik)
    z = (rng.normal(size=(C, e.D)) * 0.2).astype(F)
    _check(e, fam, z, (0, C - 1))
    keys = prng.split(prng.key(6), C)
    e.init(keys, 8)
    out = e.run(14, 8, fields=FIELDS)
    kern = chain.Kernel(device_potential(e, 2), max_tree_depth=(4, 4))
    res, _ = chain.run_chain(kern, fam, keys[2], 8, 6, fields=FIELDS)
    assert_run_equal(out, res, 2)
