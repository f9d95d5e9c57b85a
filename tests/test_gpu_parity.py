"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle.

Levels (BASELINE.json north star):
  1. PRNG streams and tree/adaptation bookkeeping: bit-exact.
  2. potential and gradient: rtol 1e-5 (fp32) against the fp64 oracle.
  3. fixed-step leapfrog trajectory: rtol 1e-4.
  4. posterior moments within 4 Monte-Carlo standard errors, split R-hat < 1.01.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng            # noqa: E402
from oracle import chain, detmath as dm, diag, families, prng      # noqa: E402

F = np.float32
FIELDS = ("z", "diverging", "num_steps", "accept_prob", "potential_energy", "energy", "step_size", "mean_accept_prob")
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0], F)
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0], F)


def eight_schools(C, **kw):
    return eng.Engine(family=_capi.FAMILY_EIGHT_SCHOOLS, num_chains=C, n_rows=8, y=Y8, aux=S8, tau_scale=5.0, **kw)


def glm_engine(C, X, y, **kw):
    return eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, **kw)


def device_potential(e, c):
    """Oracle-side potential that asks the engine's own hook (chain slot c), for bit-exact runs."""
    def pot(z):
        zz = np.zeros((e.C, e.D), F)
        zz[c] = z
        U, g = e.potential_and_grad(zz)
        return F(U[c].item()), g[c].cpu().numpy()
    return pot


def assert_run_equal(out, res, c):
    for f in FIELDS:
        got = out[f][c].cpu().numpy()
        if f == "diverging":
            got = got.astype(bool)
        np.testing.assert_array_equal(got, res[f], err_msg=f"chain {c} field {f}")


# ------------------------------------------------------------------------------------ level 1: PRNG
def test_prng_bit_exact():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2 ** 32, size=(4096, 2), dtype=np.uint64).astype(np.uint32)
    got = eng.prng_split(keys, 3)
    for i in range(0, 4096, 97):
        np.testing.assert_array_equal(got[i], prng.split(keys[i], 3))
    k = prng.key(42)
    n = 1 << 20
    np.testing.assert_array_equal(eng.prng_bits(k, n), prng.random_bits(k, n))
    np.testing.assert_array_equal(eng.prng_uniform(k, n, -2.0, 2.0), prng.uniform(k, n, -2.0, 2.0))
    m = 20000
    np.testing.assert_array_equal(eng.prng_normal(k, m), prng.normal(k, m))
    assert eng.prng_normal(k, 1)[0] == F(-0.028304616)          # value printed in JAX's docs


def test_detmath_bit_exact():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-104, 89, 3000), rng.uniform(-1, 1, 3000), [0, -0.0, np.inf, -np.inf, np.nan, 88.73, -103.98]]).astype(F)
    for op, fn in ((0, dm.exp), (3, dm.expit)):
        np.testing.assert_array_equal(eng.detmath(op, x), np.array([fn(v) for v in x], F))
    xp = np.concatenate([np.exp(rng.uniform(-100, 88, 3000)), [0, 1, np.inf, 1e-42, -1]]).astype(F)
    np.testing.assert_array_equal(eng.detmath(1, xp), np.array([dm.log(v) for v in xp], F))
    xl = np.concatenate([rng.uniform(-0.999, 5, 3000), rng.uniform(-1e-4, 1e-4, 500), [-1, 0, np.inf]]).astype(F)
    np.testing.assert_array_equal(eng.detmath(2, xl), np.array([dm.log1p(v) for v in xl], F))
    xe = np.concatenate([rng.uniform(-1, 1, 3000), [1, -1, 0, 0.99999994]]).astype(F)
    np.testing.assert_array_equal(eng.detmath(4, xe), np.array([dm.erfinv(v) for v in xe], F))


# ------------------------------------------------------------------------------------ level 2: potentials
def _check_potential(e, fam, rng, scale=0.7):
    z = (rng.normal(size=(e.C, e.D)) * scale).astype(F)
    U, g = e.potential_and_grad(z)
    U, g = U.cpu().numpy(), g.cpu().numpy()
    for c in range(e.C):
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U[c], u64, rtol=1e-5, err_msg=fam.name)
        np.testing.assert_allclose(g[c], g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max(), err_msg=fam.name)


def test_potentials_warp_regime():
    rng = np.random.default_rng(5)
    X = (rng.normal(size=(300, 6)) * 0.5).astype(F)
    yb = rng.integers(0, 2, 300).astype(F)
    yp = rng.poisson(2.0, 300).astype(F)
    yn = rng.normal(size=300).astype(F)
    e = eight_schools(3)
    assert e.regime == _capi.REGIME_WARP
    _check_potential(e, families.EightSchools(S8, Y8), rng)
    _check_potential(glm_engine(3, X, yb), families.logistic_regression(X, yb), rng)
    _check_potential(glm_engine(3, X, yp, likelihood=_capi.LIK_POISSON_LOG), families.GLM(X, yp, likelihood="poisson"), rng)
    _check_potential(glm_engine(3, X, yb, local_scales=1, global_scale=1), families.horseshoe(X, yb, "bernoulli"), rng)
    _check_potential(glm_engine(3, X, yn, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=1),
                     families.horseshoe(X, yn, "normal"), rng)
    _check_potential(glm_engine(3, X, yb, global_scale=_capi.SCALE_EXPONENTIAL, group_col_begin=2, group_col_end=5, tau_scale=2.0),
                     families.GLM(X, yb, global_scale="exponential", group_cols=(2, 5), tau_scale=2.0), rng)


@pytest.mark.parametrize("N,D,C,lik", [
    (20000, 54, 8, "bernoulli"),      # covtype-shaped, multiple tiles per CTA
    (1003, 54, 8, "bernoulli"),       # fewer row quads than CTAs, N % 4 != 0 tail rows
    (5001, 10, 3, "poisson"),         # DPL = 2, partial chain group
    (7777, 33, 5, "bernoulli"),       # DPL = 7 with odd column count (bank-conflict fallback rho)
    (4096, 64, 8, "normal_hs"),       # widest supported D, horseshoe + Normal likelihood
    (3000, 3, 2, "bernoulli"),        # reference's logistic regression test shape (test_mcmc.py:104)
])
def test_potentials_stream_regime(N, D, C, lik):
    rng = np.random.default_rng(N + D)
    X = rng.normal(size=(N, D)).astype(F)
    beta = rng.normal(size=D) * 0.3
    if lik == "bernoulli":
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ beta))).astype(F)
        e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM)
        fam = families.logistic_regression(X, y)
    elif lik == "poisson":
        X = (X * 0.3).astype(F)
        y = rng.poisson(np.exp(np.clip(X @ beta, -3, 3))).astype(F)
        e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, likelihood=_capi.LIK_POISSON_LOG)
        fam = families.GLM(X, y, likelihood="poisson")
    else:
        y = (X @ beta + 0.1 * rng.normal(size=N)).astype(F)
        e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=1)
        fam = families.horseshoe(X, y, "normal")
    assert e.regime == _capi.REGIME_STREAM
    _check_potential(e, fam, rng, scale=0.3)
    # deterministic: the same launch twice gives identical bits
    z = (rng.normal(size=(C, e.D)) * 0.3).astype(F)
    U1, g1 = e.potential_and_grad(z)
    U2, g2 = e.potential_and_grad(z)
    assert torch.equal(U1, U2) and torch.equal(g1, g2)


def test_potential_covtype_full_size():
    """BASELINE config 2 shape: N = 581012, D = 54, 8 chains (synthetic covtype-like data)."""
    rng = np.random.default_rng(1)
    N, D = 581012, 54
    X = rng.standard_normal(size=(N, D), dtype=F)
    beta = (rng.normal(size=D) * 0.3).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
    e = glm_engine(8, X, y)
    assert e.regime == _capi.REGIME_STREAM
    fam = families.logistic_regression(X, y)
    _check_potential(e, fam, rng, scale=0.2)
    # size-independent property: U(z) - U(0) == sum of per-row losses is additive over disjoint row blocks
    z = np.tile((rng.normal(size=D) * 0.2).astype(F), (8, 1))
    U_full = e.potential_and_grad(z)[0][0].item()
    half = N // 2
    e1 = glm_engine(8, X[:half], y[:half], regime=_capi.REGIME_STREAM)
    e2 = glm_engine(8, X[half:], y[half:], regime=_capi.REGIME_STREAM)
    prior = 0.5 * float(np.sum(z[0].astype(np.float64) ** 2)) + D * 0.9189385332046727
    U_sum = e1.potential_and_grad(z)[0][0].item() + e2.potential_and_grad(z)[0][0].item() - prior
    np.testing.assert_allclose(U_full, U_sum, rtol=2e-6)


# ------------------------------------------------------------------------------------ level 3: leapfrog
def test_leapfrog_trajectory():
    rng = np.random.default_rng(2)
    X = (rng.normal(size=(500, 5)) * 0.5).astype(F)
    y = rng.integers(0, 2, 500).astype(F)
    fam = families.logistic_regression(X, y)
    from oracle.tree import leapfrog
    for regime in (_capi.REGIME_WARP, _capi.REGIME_STREAM):
        e = glm_engine(2, X, y, regime=regime)
        z0 = (rng.normal(size=(2, 5)) * 0.3).astype(F)
        r0 = rng.normal(size=(2, 5)).astype(F)
        imm = np.exp(rng.normal(size=(2, 5)) * 0.2).astype(F)
        eps = np.array([0.01, -0.02], F)
        z, r, U, g = e.leapfrog(eps, imm, z0, r0, 20)
        for c in range(2):
            zz, rr = z0[c].copy(), r0[c].copy()
            u, gg = fam.potential_and_grad(zz)
            for _ in range(20):
                zz, rr, u, gg = leapfrog(fam.potential_and_grad, eps[c], imm[c], zz, rr, gg)
            np.testing.assert_allclose(z[c].cpu().numpy(), zz, rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(r[c].cpu().numpy(), rr, rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(U[c].item(), u, rtol=1e-4)


# ------------------------------------------------------------------------------------ level 1: bookkeeping
def test_warp_regime_run_bit_exact_eight_schools():
    e = eight_schools(4)
    keys = prng.split(prng.key(0), 4)
    e.init(keys, 120)
    out = e.run(180, 120, fields=FIELDS)
    fam = families.EightSchools(S8, Y8)
    for c in (0, 3):
        kern = chain.Kernel(device_potential(e, c))
        res, last = chain.run_chain(kern, fam, keys[c], 120, 60, fields=FIELDS)
        assert_run_equal(out, res, c)
    st, vec = e.state()
    np.testing.assert_array_equal(vec["inverse_mass_matrix"][3], last.adapt_state.inverse_mass_matrix)
    np.testing.assert_array_equal(np.array(st[3].rng_key), last.rng_key)
    assert st[3].i == 180 and st[3].done == 1


def test_warp_regime_run_bit_exact_options():
    """step-size heuristic + (warm-up, sampling) tree depths + thinning, then plain HMC."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(2), 2)
    e = eight_schools(2, find_heuristic_step_size=1, max_tree_depth_warmup=4, max_tree_depth=6)
    e.init(keys, 160)
    out = e.run(160 + 31, 160, thinning=3, fields=FIELDS)
    kern = chain.Kernel(device_potential(e, 1), find_heuristic_step_size=True, max_tree_depth=(4, 6))
    res, _ = chain.run_chain(kern, fam, keys[1], 160, 31, thinning=3, fields=FIELDS)
    assert_run_equal(out, res, 1)
    e = eight_schools(2, algo=_capi.ALGO_HMC, hmc_num_steps=7, step_size=0.1)
    e.init(keys, 50)
    out = e.run(80, 50, fields=FIELDS)
    kern = chain.Kernel(device_potential(e, 0), algo="HMC", num_steps=7, step_size=0.1)
    res, _ = chain.run_chain(kern, fam, keys[0], 50, 30, fields=FIELDS)
    assert_run_equal(out, res, 0)


def test_warp_regime_horseshoe_bit_exact():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(40, 5)).astype(F)
    y = (X[:, 0] * 2 - X[:, 1] + 0.05 * rng.normal(size=40)).astype(F)
    fam = families.horseshoe(X, y, "normal")
    e = glm_engine(2, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=1, max_tree_depth_warmup=6, max_tree_depth=6)
    keys = prng.split(prng.key(8), 2)
    e.init(keys, 60)
    out = e.run(80, 60, fields=FIELDS)
    kern = chain.Kernel(device_potential(e, 1), max_tree_depth=(6, 6))
    res, _ = chain.run_chain(kern, fam, keys[1], 60, 20, fields=FIELDS)
    assert_run_equal(out, res, 1)


def test_stream_regime_run_bit_exact():
    """The persistent streaming engine against the oracle driven by the engine's own potential hook."""
    rng = np.random.default_rng(4)
    N, D, C = 6000, 7, 3
    X = rng.normal(size=(N, D)).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ (rng.normal(size=D) * 0.5))))).astype(F)
    fam = families.logistic_regression(X, y)
    e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=5, max_tree_depth=5)
    keys = prng.split(prng.key(7), C)
    e.init(keys, 40)
    out = e.run(60, 40, fields=FIELDS)
    for c in (0, 2):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(5, 5))
        res, last = chain.run_chain(kern, fam, keys[c], 40, 20, fields=FIELDS)
        assert_run_equal(out, res, c)
    st, vec = e.state()
    assert all(st[c].i == 60 and st[c].done == 1 for c in range(C))
    assert sum(int(st[c].total_leapfrogs) for c in range(C)) > 0


def test_pass_bounded_launches_equal_one_unbounded_run():
    """b200nuts_run with max_passes: the chains pause anywhere in their trees and the next call continues them; the
    concatenation is bit-identical to a single unbounded run (same samples, same adaptation, same keys)."""
    rng = np.random.default_rng(12)
    N, D, C = 8000, 54, 8
    X = rng.normal(size=(N, D)).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ (rng.normal(size=D) * 0.3))))).astype(F)
    keys = prng.split(prng.key(31), C)
    a = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=6, max_tree_depth=6)
    a.init(keys, 25)
    ref = a.run(40, 25, fields=FIELDS)
    passes_ref = a.pass_count
    b = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=6, max_tree_depth=6)
    b.init(keys, 25)
    out, calls = None, 0
    while True:
        out = b.run(40, 25, fields=FIELDS, max_passes=37, out=out)
        calls += 1
        st, _ = b.state()
        if all(s.done == 1 and s.i == 40 for s in st):
            break
        assert calls < 500
    assert calls > 3 and b.pass_count == passes_ref
    for f in FIELDS:
        assert torch.equal(out[f], ref[f]), f
    sa, va = a.state()
    sb, vb = b.state()
    for c in range(C):
        assert list(sa[c].rng_key) == list(sb[c].rng_key) and sa[c].total_leapfrogs == sb[c].total_leapfrogs
    np.testing.assert_array_equal(va["inverse_mass_matrix"], vb["inverse_mass_matrix"])


def test_stream_and_warp_runs_bit_exact_with_two_elements_per_lane():
    """32 < D <= 64: every lane of the chain's warp owns two coefficients (the register-resident leaf update)."""
    rng = np.random.default_rng(77)
    N, D, C = 5000, 40, 2
    X = (rng.normal(size=(N, D)) * 0.7).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ (rng.normal(size=D) * 0.4))))).astype(F)
    fam = families.logistic_regression(X, y)
    keys = prng.split(prng.key(21), C)
    for regime in (_capi.REGIME_STREAM, _capi.REGIME_WARP):
        e = glm_engine(C, X, y, regime=regime, max_tree_depth_warmup=5, max_tree_depth=5)
        e.init(keys, 25)
        out = e.run(40, 25, fields=FIELDS)
        kern = chain.Kernel(device_potential(e, 1), max_tree_depth=(5, 5))
        res, _ = chain.run_chain(kern, fam, keys[1], 25, 15, fields=FIELDS)
        assert_run_equal(out, res, 1)


@pytest.mark.parametrize("C", [9, 20, 33])
def test_stream_regime_many_chains_rotating_groups(C):
    """More than 8 chains in the streaming engine: passes rotate over chain groups of 8 (the owners' ticks overlap with
    the other groups' sweeps).  Potentials of every chain, whole runs bit-exact against the oracle, and chain c of the
    C-chain run == the same chain of an 8-chain run (the reference's chain-independence, test_mcmc.py:868-914)."""
    rng = np.random.default_rng(40 + C)
    N, D = 7000, 11
    X = rng.normal(size=(N, D)).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ (rng.normal(size=D) * 0.5))))).astype(F)
    fam = families.logistic_regression(X, y)
    e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=5, max_tree_depth=5)
    z = (rng.normal(size=(C, D)) * 0.3).astype(F)
    U, g = e.potential_and_grad(z)
    for c in range(C):
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U[c].item(), u64, rtol=1e-5)
        np.testing.assert_allclose(g[c].cpu().numpy(), g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max())
    keys = prng.split(prng.key(9), C)
    e.init(keys, 30)
    out = e.run(45, 30, fields=FIELDS)
    for c in (0, 8, C - 1):
        kern = chain.Kernel(device_potential(e, c), max_tree_depth=(5, 5))
        res, _ = chain.run_chain(kern, fam, keys[c], 30, 15, fields=FIELDS)
        assert_run_equal(out, res, c)
    e8 = glm_engine(8, X, y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=5, max_tree_depth=5)
    e8.init(keys[C - 8:], 30)
    out8 = e8.run(45, 30, fields=FIELDS)
    for f in FIELDS:
        assert torch.equal(out[f][C - 8:], out8[f]), f
    st, _ = e.state()
    assert all(st[c].i == 45 and st[c].done == 1 for c in range(C))


def test_chain_of_many_equals_single_chain_and_resume():
    """test/infer/test_mcmc.py:868-914 (chain 0 of a 2-chain run == 1-chain run with split(key)[0])
    and :437-485 (warmup then run == run)."""
    keys = prng.split(prng.key(3), 2)
    e2 = eight_schools(2)
    e2.init(keys, 80)
    two = e2.run(120, 80, fields=FIELDS)
    e1 = eight_schools(1)
    e1.init(keys[:1], 80)
    e1.run(80, 80, fields=FIELDS)                       # warmup only
    st, _ = e1.state()
    assert st[0].i == 80
    one = e1.run(120, 80, fields=FIELDS)                # resume
    for f in FIELDS:
        assert torch.equal(two[f][0], one[f][0]), f


def test_state_roundtrip_resume():
    """post_warmup_state style resume through get_state / set_state (mcmc.py:558-587)."""
    keys = prng.split(prng.key(5), 2)
    a = eight_schools(2)
    a.init(keys, 60)
    a.run(60, 60)
    st, vec = a.state()
    ref = a.run(90, 60, fields=FIELDS)
    b = eight_schools(2)
    b.set_state(st, vec, 60)
    got = b.run(90, 60, fields=FIELDS)
    for f in FIELDS:
        assert torch.equal(ref[f], got[f]), f


# ------------------------------------------------------------------------------------ level 4: posteriors
def test_eight_schools_posterior_matches_readme_and_oracle():
    """BASELINE config 1: 4 chains, 1000 warm-up / 1000 samples.  README.md:118-143 prints (1 chain, 500/1000)
    mu 4.08 +- 3.51 (n_eff 720), tau 3.96 +- 3.31 (n_eff 489), theta[0] 6.48 +- 5.72, 0 divergences, all r_hat 1.00,
    E[log joint] -46.09.  The engine's chain 0 is compared with the ORACLE's whole run of the same key (2000 transitions,
    bit-exact), the pooled posterior with the README table within 4 combined Monte-Carlo standard errors."""
    e = eight_schools(4)
    keys = prng.split(prng.key(0), 4)
    e.init(keys, 1000)
    out = e.run(2000, 1000, fields=FIELDS)
    res, _ = chain.run_chain(chain.Kernel(device_potential(e, 0)), families.EightSchools(S8, Y8), keys[0], 1000, 1000, fields=FIELDS)
    assert_run_equal(out, res, 0)
    con = e.constrain(out["z"]).view(4, 1000, -1).cpu().numpy().astype(np.float64)
    mu, tau, theta = con[..., 0], con[..., 1], con[..., 10:18]
    assert int(out["diverging"].sum().item()) <= 5
    for x, (mean, sd, n_eff) in ((mu, (4.08, 3.51, 720)), (tau, (3.96, 3.31, 489)), (theta[..., 0], (6.48, 5.72, 802))):
        ess = diag.effective_sample_size(x)
        mcse = np.sqrt(x.std() ** 2 / ess + sd ** 2 / n_eff)         # ours and the README's
        assert abs(x.mean() - mean) < 4 * mcse, (x.mean(), mean, mcse)
        assert diag.split_gelman_rubin(x) < 1.01
    assert abs(-out["potential_energy"].mean().item() - (-46.09)) < 0.5
    assert 3 < out["num_steps"].float().mean().item() < 15


@pytest.mark.parametrize("case", ["poisson_heuristic", "horseshoe_normal_thinning", "bernoulli_hmc"])
def test_stream_regime_runs_all_likelihoods_bit_exact(case):
    """Whole runs in the STREAMING regime (persistent kernel, tagged exchange) for the paths the round-1 tests only covered
    in the warp regime: find_heuristic_step_size (hmc_util.py:314-384), Poisson and Normal likelihoods, local + global
    scales (horseshoe), thinning (util.py:375,388), plain HMC -- bit-exact against the oracle."""
    rng = np.random.default_rng(31)
    N, D, C = 20000, 20, 5
    X = (rng.normal(size=(N, D)) * 0.5).astype(F)
    beta = rng.normal(size=D) * 0.3
    kw, okw, thin = {}, {}, 1
    if case == "poisson_heuristic":
        y = rng.poisson(np.exp(np.clip(X @ beta, -4, 4))).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_POISSON_LOG, regime=_capi.REGIME_STREAM, find_heuristic_step_size=1,
                       max_tree_depth_warmup=6, max_tree_depth=6)
        fam = families.GLM(X, y, likelihood="poisson")
        okw = dict(find_heuristic_step_size=True, max_tree_depth=(6, 6))
    elif case == "horseshoe_normal_thinning":
        y = (X @ beta + 0.1 * rng.normal(size=N)).astype(F)
        e = glm_engine(C, X, y, likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                       regime=_capi.REGIME_STREAM, max_tree_depth_warmup=6, max_tree_depth=6)
        fam = families.horseshoe(X, y, "normal")
        okw, thin = dict(max_tree_depth=(6, 6)), 3
    else:
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
        e = glm_engine(C, X, y, regime=_capi.REGIME_STREAM, algo=_capi.ALGO_HMC, hmc_num_steps=6, step_size=0.01)
        fam = families.logistic_regression(X, y)
        okw = dict(algo="HMC", num_steps=6, step_size=0.01)
    assert e.regime == _capi.REGIME_STREAM
    keys = prng.split(prng.key(32), C)
    e.init(keys, 30)
    out = e.run(48, 30, thinning=thin, fields=FIELDS)
    for c in (0, C - 1):
        res, _ = chain.run_chain(chain.Kernel(device_potential(e, c), **okw), fam, keys[c], 30, 18, thinning=thin, fields=FIELDS)
        assert_run_equal(out, res, c)


def test_logistic_regression_posterior():
    """test/infer/test_mcmc.py:104-168: N = 3000, 3 coefficients, posterior mean within 0.4... here
    within 4 MCSE of the oracle-independent truth check and split R-hat < 1.01, both regimes."""
    rng = np.random.default_rng(0)
    N, D = 3000, 3
    X = rng.normal(size=(N, D)).astype(F)
    true = np.array([1.0, 2.0, 3.0])
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ true)))).astype(F)
    means = []
    for regime in (_capi.REGIME_WARP, _capi.REGIME_STREAM):
        e = glm_engine(4, X, y, regime=regime)
        e.init(prng.split(prng.key(1), 4), 500)
        out = e.run(1000, 500, fields=("z", "diverging"))
        z = out["z"].cpu().numpy().astype(np.float64)
        np.testing.assert_allclose(z.mean(axis=(0, 1)), true, atol=0.4)
        assert np.all(diag.split_gelman_rubin(z) < 1.01)
        means.append((z.mean(axis=(0, 1)), z.std(axis=(0, 1)), diag.effective_sample_size(z)))
    (m1, s1, n1), (m2, s2, n2) = means
    assert np.all(np.abs(m1 - m2) < 4 * np.sqrt(s1 ** 2 / n1 + s2 ** 2 / n2))
