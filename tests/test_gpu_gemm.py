"""Many-chain GEMM regime (R3: tcgen05 + TMEM, csrc/gemm_engine.cuh) against the CPU oracle, through the C ABI.

Same three levels as the other regimes (BASELINE.json north star): potential / gradient rtol 1e-5 against the fp64
oracle; whole runs (tree bookkeeping, adaptation, PRNG) bit-exact against the oracle driven by the engine's own potential
hook; and the BASELINE shapes of configs 3 (hierarchical GLM, N = 100k, D = 256, 16384 chains) and 4 (horseshoe,
N = 10k, D = 1000, 1024 chains) in this regime.  Reference path: vmapped ``sample_fn`` numpyro/infer/hmc.py:790-798,
models examples/horseshoe_regression.py:37-78.
"""
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


def _check(e, fam, z, chains, rtol=1e-5):
    U, g = e.potential_and_grad(z)
    U, g = U.cpu().numpy(), g.cpu().numpy()
    for c in chains:
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U[c], u64, rtol=rtol, err_msg=f"chain {c}")
        np.testing.assert_allclose(g[c], g64, rtol=rtol, atol=rtol * np.abs(g64).max(), err_msg=f"chain {c}")
    return U, g


def _bernoulli(rng, X, beta):
    return (rng.uniform(size=X.shape[0]) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)


@pytest.mark.parametrize("lik", ["bernoulli", "poisson", "normal"])
def test_gemm_potential_small(lik):
    """Plain / horseshoe GLMs with a ragged shape: N not a multiple of 128, D not a multiple of 32, C not a multiple of 128."""
    rng = np.random.default_rng(7)
    N, D, C = 1000, 40, 200
    X = (rng.normal(size=(N, D)) * 0.5).astype(F)
    beta = rng.normal(size=D) * 0.4
    if lik == "bernoulli":
        y = _bernoulli(rng, X, beta)
        e = glm_engine(C, X, y, regime=_capi.REGIME_GEMM)
        fam = families.logistic_regression(X, y)
    elif lik == "poisson":
        y = rng.poisson(np.exp(np.clip(X @ beta, -5, 5))).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_POISSON_LOG, regime=_capi.REGIME_GEMM)
        fam = families.GLM(X, y, likelihood="poisson")
    else:
        y = (X @ beta + 0.1 * rng.normal(size=N)).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY, regime=_capi.REGIME_GEMM)
        fam = families.horseshoe(X, y, "normal")
    assert e.regime == _capi.REGIME_GEMM
    info = e.gemm_info()
    assert info["chain_tiles"] == 2 and info["row_chunks"] == 8 and info["k_blocks"] == 2 and info["column_blocks"] == 1
    z = (rng.normal(size=(C, e.D)) * 0.3).astype(F)
    U, g = _check(e, fam, z, (0, 1, 77, 127, 128, C - 1))
    # the hook is deterministic, and a chain's result does not depend on what the other chains evaluate
    U2, g2 = e.potential_and_grad(z)
    assert torch.equal(torch.as_tensor(U), U2.cpu()) and torch.equal(torch.as_tensor(g), g2.cpu())
    z3 = np.zeros_like(z)
    z3[77] = z[77]
    U3, g3 = e.potential_and_grad(z3)
    assert U3[77].item() == U[77] and np.array_equal(g3[77].cpu().numpy(), g[77])


def test_gemm_potential_column_blocks():
    """More than 256 columns: the backward product runs per column block of 256 (here 256 + 64 of 300 columns)."""
    rng = np.random.default_rng(8)
    N, D, C = 700, 300, 130
    X = (rng.normal(size=(N, D)) / np.sqrt(D)).astype(F)
    y = _bernoulli(rng, X, rng.normal(size=D))
    e = glm_engine(C, X, y, regime=_capi.REGIME_GEMM)
    info = e.gemm_info()
    assert info["column_blocks"] == 2 and info["padded_columns"] == 320
    z = (rng.normal(size=(C, D)) * 0.5).astype(F)
    _check(e, families.logistic_regression(X, y), z, (0, 64, 129))


def test_gemm_run_bit_exact_small():
    """Whole NUTS runs (init retries, adaptation windows, trees, thinning) bit-exact against the oracle driven by the
    engine's potential hook; chains of both tiles, a segment count that does not divide the chunk count."""
    rng = np.random.default_rng(9)
    N, D, C = 1500, 24, 160
    X = rng.normal(size=(N, D)).astype(F)
    y = _bernoulli(rng, X, rng.normal(size=D) * 0.5)
    e = glm_engine(C, X, y, regime=_capi.REGIME_GEMM, max_tree_depth_warmup=6, max_tree_depth=6)
    fam = families.logistic_regression(X, y)
    keys = prng.split(prng.key(11), C)
    e.init(keys, 30)
    out = e.run(45, 30, fields=FIELDS)
    assert e.pass_count > 45
    for c in (3, 159):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(6, 6))
        res, _ = chain.run_chain(kern, fam, keys[c], 30, 15, fields=FIELDS)
        assert_run_equal(out, res, c)
    assert int(out["num_steps"].sum().item()) >= C * 15
    # pass-bounded launches: chains pause between passes and resume bit-identically
    e2 = glm_engine(C, X, y, regime=_capi.REGIME_GEMM, max_tree_depth_warmup=6, max_tree_depth=6)
    e2.init(keys, 30)
    out2, calls = None, 0
    while True:
        out2 = e2.run(45, 30, fields=FIELDS, max_passes=37, out=out2)
        calls += 1
        st, _ = e2.state()
        if all(s.done for s in st):
            break
        assert calls < 200
    assert calls > 3
    for f in FIELDS:
        assert torch.equal(out[f], out2[f]), f


def test_gemm_horseshoe_heuristic_step_size_and_hmc():
    rng = np.random.default_rng(10)
    N, D, C = 400, 33, 128
    X = rng.normal(size=(N, D)).astype(F)
    X -= X.mean(0)
    y = (2 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2] + 0.05 * rng.normal(size=N)).astype(F)
    e = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                   regime=_capi.REGIME_GEMM, max_tree_depth_warmup=5, max_tree_depth=5, find_heuristic_step_size=1)
    fam = families.horseshoe(X, y, "normal")
    assert e.D == 2 * D + 2
    keys = prng.split(prng.key(12), C)
    e.init(keys, 25)
    out = e.run(33, 25, thinning=2, fields=FIELDS)
    kern = chain.Kernel(device_potential(e, 100), max_tree_depth=(5, 5), find_heuristic_step_size=True)
    res, _ = chain.run_chain(kern, fam, keys[100], 25, 8, thinning=2, fields=FIELDS)
    assert_run_equal(out, res, 100)
    # plain HMC on the same engine type
    h = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                   regime=_capi.REGIME_GEMM, algo=_capi.ALGO_HMC, hmc_num_steps=5, step_size=0.01)
    h.init(keys, 10)
    outh = h.run(16, 10, fields=FIELDS)
    kern = chain.Kernel(device_potential(h, 5), algo="HMC", num_steps=5, step_size=0.01)
    res, _ = chain.run_chain(kern, fam, keys[5], 10, 6, fields=FIELDS)
    assert_run_equal(outh, res, 5)


def test_auto_regime_picks_gemm_for_many_chains():
    rng = np.random.default_rng(13)
    X = rng.normal(size=(2048, 64)).astype(F)
    y = _bernoulli(rng, X, rng.normal(size=64) * 0.3)
    assert glm_engine(256, X, y).regime == _capi.REGIME_GEMM
    assert glm_engine(8, X, y).regime != _capi.REGIME_GEMM


# --------------------------------------------------------------------------------------- BASELINE shapes
def config3_data(lik="bernoulli", N=100_000, D=256, seed=33):
    """BASELINE config 3: X ~ N(0,1)/sqrt(D) with a one-hot block of 64 group columns whose
    coefficients share a global scale (non-centred random intercepts)."""
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal(size=(N, D), dtype=np.float32) / np.sqrt(D)).astype(F)
    X[:, 192:] = 0.0
    X[np.arange(N), 192 + rng.integers(0, 64, size=N)] = 1.0
    beta = rng.normal(size=D) * 0.5
    eta = np.clip(X @ beta, -10, 10)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(F) if lik == "bernoulli" else rng.poisson(np.exp(eta)).astype(F)
    return X, y


@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_config3_full_shape_gemm(lik):
    """N = 100k, D = 256, 16384 chains, hierarchical GLM."""
    N, D, C = 100_000, 256, 16384
    X, y = config3_data(lik)
    kw = dict(likelihood=_capi.LIK_POISSON_LOG) if lik == "poisson" else {}
    e = glm_engine(C, X, y, global_scale=_capi.SCALE_HALFCAUCHY, group_col_begin=192, group_col_end=256, tau_scale=1.0,
                   max_tree_depth_warmup=5, max_tree_depth=5, **kw)
    assert e.regime == _capi.REGIME_GEMM and e.D == D + 1
    info = e.gemm_info()
    assert info["chain_tiles"] == 128 and info["row_chunks"] == 782 and info["k_blocks"] == 8
    fam = families.GLM(X, y, global_scale="halfcauchy", group_cols=(192, 256), tau_scale=1.0,
                       **(dict(likelihood="poisson") if lik == "poisson" else {}))
    rng = np.random.default_rng(3)
    z = (rng.normal(size=(C, e.D)) * 0.3).astype(F)
    _check(e, fam, z, (0, 5000, 12345, C - 1))
    if lik == "poisson":
        return                                   # (the whole-run comparison below is likelihood independent)
    keys = prng.split(prng.key(5), C)
    e.init(keys, 6)
    out = e.run(10, 6, fields=FIELDS)
    for c in (7, C - 1):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(5, 5))
        res, _ = chain.run_chain(kern, fam, keys[c], 6, 4, fields=FIELDS)
        assert_run_equal(out, res, c)
    assert torch.isfinite(out["z"]).all() and int(out["num_steps"].sum().item()) >= 4 * C
    # deeper trees (fixed small step, up to 31 leapfrogs per transition), all 16384 chains asynchronous
    d = glm_engine(C, X, y, global_scale=_capi.SCALE_HALFCAUCHY, group_col_begin=192, group_col_end=256, tau_scale=1.0,
                   max_tree_depth_warmup=5, max_tree_depth=5, step_size=0.02, adapt_step_size=0)
    d.init(keys, 2)
    outd = d.run(5, 2, fields=FIELDS)
    assert outd["num_steps"].float().mean().item() > 4 and int(outd["num_steps"].max().item()) == 31
    for c in (1, int(outd["num_steps"].sum(dim=1).argmax().item())):
        kern = chain.Kernel(device_potential(d, c), max_tree_depth=(5, 5), step_size=0.02, adapt_step_size=False)
        res, _ = chain.run_chain(kern, fam, keys[c], 2, 3, fields=FIELDS)
        assert_run_equal(outd, res, c)


@pytest.mark.parametrize("lik", ["normal", "bernoulli"])
def test_config4_full_shape_gemm(lik):
    """N = 10k, D = 1000, 1024 chains, horseshoe regression (examples/horseshoe_regression.py:105-125 data recipe)."""
    N, D, C = 10_000, 1000, 1024
    rng = np.random.default_rng(0)
    X = rng.standard_normal(size=(N, D), dtype=np.float32)
    X -= X.mean(0)
    eta = 2 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2]
    if lik == "normal":
        y = (eta + 0.05 * rng.normal(size=N)).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                       max_tree_depth_warmup=4, max_tree_depth=4)
        assert e.D == 2 * D + 2
    else:
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(F)
        e = glm_engine(C, X, y, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY, max_tree_depth_warmup=4, max_tree_depth=4)
        assert e.D == 2 * D + 1
    assert e.regime == _capi.REGIME_GEMM
    info = e.gemm_info()
    assert info["chain_tiles"] == 8 and info["column_blocks"] == 4 and info["padded_columns"] == 1024
    fam = families.horseshoe(X, y, lik)
    z = (rng.normal(size=(C, e.D)) * 0.2).astype(F)
    _check(e, fam, z, (0, 500, C - 1))
    keys = prng.split(prng.key(6), C)
    e.init(keys, 5)
    out = e.run(8, 5, fields=FIELDS)
    for c in (2, C - 1):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(4, 4))
        res, _ = chain.run_chain(kern, fam, keys[c], 5, 3, fields=FIELDS)
        assert_run_equal(out, res, c)
