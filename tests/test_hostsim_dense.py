"""Dense / user-supplied mass matrices (numpyro/infer/hmc_util.py:192-237, 439-515, 726-728,
1193-1194; hmc.py:759-769): the engine's state machine (tick.cuh, host build) against the NumPy oracle, bit for bit, plus
the linear algebra of the oracle against NumPy's own factorisations."""
import numpy as np
import pytest

import hostsim_util as hs
from numpyro_b200 import _capi
from oracle import adapt, chain, families, prng, tree

F = np.float32
FIELDS = ("z", "diverging", "num_steps", "accept_prob", "potential_energy", "energy", "step_size", "mean_accept_prob")
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0], F)
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0], F)


def _spd(d, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(d, d))
    return (a @ a.T / d + 0.5 * np.eye(d)).astype(F)


def test_oracle_roots_against_numpy():
    """tril_inv^T tril_inv = M^-1 and sqrt sqrt^T = M (hmc_util.py:228-233 is distributions.util.cholesky_of_inverse)."""
    for d in (1, 3, 10, 37):
        imm = _spd(d, d)
        np.testing.assert_allclose(adapt.cholesky_lower(imm), np.linalg.cholesky(imm.astype(np.float64)), rtol=2e-5, atol=2e-6)
        sqrt_m, sqrt_inv = adapt.mass_matrix_roots(imm)
        assert np.allclose(np.triu(sqrt_m, 1), 0) and np.allclose(np.triu(sqrt_inv, 1), 0)         # both lower triangular
        np.testing.assert_allclose(sqrt_inv.astype(np.float64).T @ sqrt_inv, imm, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sqrt_m.astype(np.float64) @ sqrt_m.T, np.linalg.inv(imm.astype(np.float64)), rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(sqrt_m.astype(np.float64) @ sqrt_inv, np.eye(d), atol=1e-4)
    assert np.isnan(adapt.cholesky_lower(np.array([[1.0, 2.0], [2.0, 1.0]], F))).any()              # not positive definite
    # the dense product reduces to the diagonal one for a diagonal matrix
    r = np.random.default_rng(0).normal(size=9).astype(F)
    dg = np.exp(np.random.default_rng(1).normal(size=9)).astype(F)
    np.testing.assert_array_equal(tree.imm_apply(np.diag(dg), r), tree.imm_apply(dg, r))
    # Welford with outer products: cov of the samples (hmc_util.py:178-195, 213-214)
    xs = np.random.default_rng(2).normal(size=(40, 4)).astype(F) @ _spd(4, 3)
    st = adapt.welford_init(4, dense=True)
    for x in xs:
        st = adapt.welford_update(x, st)
    cov, sqrt_m, sqrt_inv = adapt.welford_final(st, regularize=False)
    np.testing.assert_allclose(cov, np.cov(xs.astype(np.float64), rowvar=False), rtol=1e-4, atol=1e-5)


def _cfg8(C, **kw):
    return _capi.default_config(family=_capi.FAMILY_EIGHT_SCHOOLS, num_chains=C, n_rows=8, y=hs._p(Y8), aux=hs._p(S8), tau_scale=5.0, **kw)


def _compare(out, fam, keys, W, S, imm=None, **kw):
    last = None
    for c in range(len(keys)):
        kern = chain.Kernel(fam.potential_and_grad, **kw)
        res, last = chain.run_chain(kern, fam, keys[c], W, S, fields=FIELDS, inverse_mass_matrix=imm)
        for f in FIELDS:
            got = out[f][c].astype(bool) if f == "diverging" else out[f][c]
            np.testing.assert_array_equal(got, res[f], err_msg=f"chain {c} field {f}")
    return last


def test_dense_mass_nuts_with_adaptation_bit_exact():
    """test_mcmc.py:75-101 (dense_mass=True): whole runs incl. the Welford outer products, the Cholesky of the reversed
    covariance and the triangular solve at every window end."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(3), 2)
    sim = hs.HostSim(_cfg8(2, dense_mass=1))
    sim.init(keys, 150)
    out = sim.run(190, 150, potential=lambda c, z: fam.potential_and_grad(z))
    last = _compare(out, fam, keys, 150, 40, dense_mass=True)
    ds = sim.dense_state()
    a = last.adapt_state
    assert a.inverse_mass_matrix.shape == (10, 10) and not np.allclose(a.inverse_mass_matrix, np.diag(np.diag(a.inverse_mass_matrix)))
    np.testing.assert_array_equal(ds["inverse_mass_matrix"][1], a.inverse_mass_matrix)
    np.testing.assert_array_equal(ds["mass_matrix_sqrt"][1], a.mass_matrix_sqrt)
    np.testing.assert_array_equal(ds["mass_matrix_sqrt_inv"][1], a.mass_matrix_sqrt_inv)
    np.testing.assert_array_equal(ds["wf_m2"][1], a.mm_state.m2)


def test_dense_mass_with_lookahead_heuristic_and_hmc():
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(8), 2)
    sim = hs.HostSim(_cfg8(2, dense_mass=1, find_heuristic_step_size=1, max_tree_depth_warmup=5, max_tree_depth=6))
    sim.set_lookahead(True)
    sim.init(keys, 160)
    out = sim.run(180, 160, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 160, 20, dense_mass=True, find_heuristic_step_size=True, max_tree_depth=(5, 6))
    sim = hs.HostSim(_cfg8(2, dense_mass=1, algo=_capi.ALGO_HMC, hmc_num_steps=5, step_size=0.1))
    sim.init(keys, 100)
    out = sim.run(120, 100, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 100, 20, dense_mass=True, algo="HMC", num_steps=5, step_size=0.1)


@pytest.mark.parametrize("dense", [False, True])
def test_user_supplied_inverse_mass_matrix(dense):
    """hmc_util.py:494-515: a supplied matrix is used as given (its roots are derived from it); with
    adapt_mass_matrix=False it never changes (test_mcmc.py:312-345 style)."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(11), 2)
    imm = _spd(10, 5) if dense else np.exp(np.random.default_rng(5).normal(size=10) * 0.3).astype(F)
    sim = hs.HostSim(_cfg8(2, dense_mass=int(dense), adapt_mass_matrix=0))
    sim.set_inverse_mass_matrix(imm)
    sim.init(keys, 60)
    out = sim.run(80, 60, potential=lambda c, z: fam.potential_and_grad(z))
    last = _compare(out, fam, keys, 60, 20, imm=imm, dense_mass=dense, adapt_mass_matrix=False)
    if dense:
        ds = sim.dense_state()
        np.testing.assert_array_equal(ds["inverse_mass_matrix"][0], imm)
        np.testing.assert_array_equal(ds["mass_matrix_sqrt"][0], last.adapt_state.mass_matrix_sqrt)
    else:
        st, z, g, im, sm = sim.state()
        np.testing.assert_array_equal(im[0], imm)
        np.testing.assert_array_equal(sm[0], last.adapt_state.mass_matrix_sqrt)
    # a 1-D matrix on a dense handle goes on the diagonal, a 2-D matrix on a diagonal handle keeps its diagonal (:496-497, :511-512)
    other = np.exp(np.random.default_rng(6).normal(size=10) * 0.3).astype(F) if dense else _spd(10, 6)
    sim = hs.HostSim(_cfg8(1, dense_mass=int(dense), adapt_mass_matrix=0))
    sim.set_inverse_mass_matrix(other)
    sim.init(keys[:1], 20)
    out = sim.run(30, 20, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys[:1], 20, 10, imm=other, dense_mass=dense, adapt_mass_matrix=False)


def test_dense_mass_wide_vector_two_rows_per_lane():
    rng = np.random.default_rng(0)
    D = 40
    mu = rng.normal(size=D).astype(F)
    sg = np.exp(rng.normal(size=D) * 0.5).astype(F)
    aux = np.concatenate([mu, sg])
    fam = families.DiagGaussian(mu, sg)
    keys = prng.split(prng.key(9), 1)
    cfg = _capi.default_config(family=_capi.FAMILY_DIAG_GAUSSIAN, num_chains=1, n_rows=D, aux=hs._p(aux), dense_mass=1)
    sim = hs.HostSim(cfg, keep=[aux])
    sim.init(keys, 110)
    out = sim.run(120, 110, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 110, 10, dense_mass=True)
