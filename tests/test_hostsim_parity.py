"""Bit-exact parity of the engine's per-chain state machine (numpyro_b200/csrc/tick.cuh, compiled
for the host by tests/hostsim) against the NumPy oracle, on CPU.

The potential is supplied by the oracle through a callback, so every remaining difference would be
in PRNG streams, leapfrog arithmetic, tree bookkeeping, adaptation or collection -- all of which must
match to the last bit (north-star level 1).  The same comparisons run on the GPU in
tests/test_gpu_parity.py with the CUDA build of the same source."""
import numpy as np
import pytest

import hostsim_util as hs
from numpyro_b200 import _capi
from oracle import chain, families, prng

FIELDS = ("z", "diverging", "num_steps", "accept_prob", "potential_energy", "energy", "step_size",
          "mean_accept_prob")
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0], np.float32)
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0], np.float32)


def _compare(sim_out, fam, keys, num_warmup, num_samples, thinning=1, z0=None, **kernel_kw):
    for c in range(len(keys)):
        kern = chain.Kernel(fam.potential_and_grad, **kernel_kw)
        res, last = chain.run_chain(kern, fam, keys[c], num_warmup, num_samples, thinning=thinning,
                                    init_z=None if z0 is None else z0[c], fields=FIELDS)
        for f in FIELDS:
            want = res[f]
            got = sim_out[f][c]
            if f == "diverging":
                got = got.astype(bool)
            np.testing.assert_array_equal(got, want, err_msg=f"chain {c} field {f}")
    return last


def _eight_schools_cfg(C, **kw):
    return _capi.default_config(family=_capi.FAMILY_EIGHT_SCHOOLS, num_chains=C, n_rows=8, y=hs._p(Y8),
                                aux=hs._p(S8), tau_scale=5.0, **kw)


def test_eight_schools_nuts_with_adaptation():
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(0), 2)
    sim = hs.HostSim(_eight_schools_cfg(2))
    sim.init(keys, 150)
    out = sim.run(220, 150, potential=lambda c, z: fam.potential_and_grad(z))
    last = _compare(out, fam, keys, 150, 70)
    st, z, g, imm, sm = sim.state()
    np.testing.assert_array_equal(imm[1], last.adapt_state.inverse_mass_matrix)
    np.testing.assert_array_equal(sm[1], last.adapt_state.mass_matrix_sqrt)
    np.testing.assert_array_equal(np.array(st[1].rng_key), last.rng_key)
    np.testing.assert_array_equal(np.array(st[1].adapt_rng_key), last.adapt_state.rng_key)
    assert st[1].i == 220 and st[1].window_idx == last.adapt_state.window_idx


def test_thinning_and_warmup_collection():
    fam = families.EightSchools(S8, Y8)
    keys = prng.key(5)[None]
    sim = hs.HostSim(_eight_schools_cfg(1))
    sim.init(keys, 30)
    out = sim.run(30 + 41, 30, thinning=3, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 30, 41, thinning=3)


def test_heuristic_step_size_and_tree_depth_pair():
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(2), 2)
    sim = hs.HostSim(_eight_schools_cfg(2, find_heuristic_step_size=1, max_tree_depth_warmup=4, max_tree_depth=6))
    sim.init(keys, 160)
    out = sim.run(190, 160, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 160, 30, find_heuristic_step_size=True, max_tree_depth=(4, 6))


def test_wide_gaussian_more_than_one_lane_row():
    rng = np.random.default_rng(0)
    D = 70
    mu = rng.normal(size=D).astype(np.float32)
    sg = np.exp(rng.normal(size=D)).astype(np.float32)
    aux = np.concatenate([mu, sg])
    fam = families.DiagGaussian(mu, sg)
    keys = prng.split(prng.key(9), 2)
    cfg = _capi.default_config(family=_capi.FAMILY_DIAG_GAUSSIAN, num_chains=2, n_rows=D, aux=hs._p(aux))
    sim = hs.HostSim(cfg, keep=[aux])
    sim.init(keys, 120)
    out = sim.run(150, 120, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 120, 30)


def test_raw_potential_key_path_and_given_init():
    """model_built=0 (ndarray mass matrix: no extra split) + init_params given."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(4), 2)
    z0 = np.random.default_rng(1).normal(size=(2, 10)).astype(np.float32) * 0.3
    sim = hs.HostSim(_eight_schools_cfg(2, model_built=0))
    sim.init(keys, 40, z0=z0)
    out = sim.run(60, 40, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 40, 20, z0=z0, model_built=False)


@pytest.mark.parametrize("num_steps", [7, 0])
def test_plain_hmc(num_steps):
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(6), 2)
    kw = dict(algo=_capi.ALGO_HMC, hmc_num_steps=num_steps, step_size=0.1)
    okw = dict(algo="HMC", num_steps=num_steps or None, step_size=0.1)
    if num_steps == 0:
        kw["trajectory_length"] = 1.5
        okw["trajectory_length"] = 1.5
    sim = hs.HostSim(_eight_schools_cfg(2, **kw))
    sim.init(keys, 60)
    out = sim.run(90, 60, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 60, 30, **okw)


def test_horseshoe_trace_order_init():
    """Init draws follow model-trace order (lambdas, tau, unscaled_betas, prec_obs) while the flat
    layout is sorted by name (infer/util.py:454-463, hmc.py:765-768)."""
    rng = np.random.default_rng(3)
    X = rng.normal(size=(30, 5)).astype(np.float32)
    y = (X[:, 0] * 2 - X[:, 1] + 0.05 * rng.normal(size=30)).astype(np.float32)
    fam = families.horseshoe(X, y, "normal")
    keys = prng.split(prng.key(8), 2)
    cfg = _capi.default_config(family=_capi.FAMILY_GLM, num_chains=2, n_rows=30, n_cols=5, X=hs._p(X), y=hs._p(y),
                               likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY,
                               max_tree_depth_warmup=6, max_tree_depth=6)
    sim = hs.HostSim(cfg, keep=[X, y])
    assert sim.D == 12
    sim.init(keys, 50)
    out = sim.run(65, 50, potential=lambda c, z: fam.potential_and_grad(z))
    _compare(out, fam, keys, 50, 15, max_tree_depth=(6, 6))


def test_warmup_then_run_equals_single_run():
    """test/infer/test_mcmc.py:437-485: warmup() followed by run() continues the same chain."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.key(11)[None]
    pot = lambda c, z: fam.potential_and_grad(z)
    a = hs.HostSim(_eight_schools_cfg(1))
    a.init(keys, 50)
    one = a.run(80, 50, potential=pot)
    b = hs.HostSim(_eight_schools_cfg(1))
    b.init(keys, 50)
    b.run(50, 50, potential=pot)
    st, *_ = b.state()
    assert st[0].i == 50 and st[0].done == 1
    two = b.run(80, 50, potential=pot)
    for f in FIELDS:
        np.testing.assert_array_equal(one[f], two[f])


def test_host_build_of_family_potentials_matches_oracle():
    """The in-warp potentials (families.cuh) against the fp64 oracle, rtol 1e-5 (north-star level 2)."""
    rng = np.random.default_rng(5)
    X = (rng.normal(size=(200, 6)) * 0.5).astype(np.float32)
    cases = [
        (families.EightSchools(S8, Y8), _eight_schools_cfg(3), []),
    ]
    yb = rng.integers(0, 2, 200).astype(np.float32)
    yp = rng.poisson(2.0, 200).astype(np.float32)
    yn = rng.normal(size=200).astype(np.float32)
    glm = lambda **kw: _capi.default_config(family=_capi.FAMILY_GLM, num_chains=3, n_rows=200, n_cols=6, X=hs._p(X), **kw)
    cases += [
        (families.logistic_regression(X, yb), glm(y=hs._p(yb)), [yb]),
        (families.GLM(X, yp, likelihood="poisson"), glm(y=hs._p(yp), likelihood=_capi.LIK_POISSON_LOG), [yp]),
        (families.horseshoe(X, yb, "bernoulli"), glm(y=hs._p(yb), local_scales=1, global_scale=1), [yb]),
        (families.horseshoe(X, yn, "normal"), glm(y=hs._p(yn), likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=1), [yn]),
        (families.GLM(X, yb, global_scale="exponential", group_cols=(2, 5), tau_scale=2.0),
         glm(y=hs._p(yb), global_scale=_capi.SCALE_EXPONENTIAL, group_col_begin=2, group_col_end=5, tau_scale=2.0), [yb]),
    ]
    for fam, cfg, keep in cases:
        sim = hs.HostSim(cfg, keep=[X] + keep)
        assert sim.D == fam.dim, fam.name
        z = (rng.normal(size=(3, fam.dim)) * 0.7).astype(np.float32)
        U, g = sim.potential(z)
        for c in range(3):
            u64, g64 = fam.potential64(z[c].astype(np.float64))
            np.testing.assert_allclose(U[c], u64, rtol=1e-5, err_msg=fam.name)
            np.testing.assert_allclose(g[c], g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max(), err_msg=fam.name)


@pytest.mark.parametrize("heuristic", [0, 1])
def test_prng_lookahead_changes_nothing(heuristic):
    """Tick::prefetch (the PRNG look-ahead the streaming engine runs while it sweeps X) must be invisible:
    same samples, same keys, same adaptation, bit for bit, against the oracle."""
    fam = families.EightSchools(S8, Y8)
    keys = prng.split(prng.key(11), 2)
    sim = hs.HostSim(_eight_schools_cfg(2, find_heuristic_step_size=heuristic, max_tree_depth_warmup=5, max_tree_depth=7))
    sim.set_lookahead(True)
    sim.init(keys, 160)
    out = sim.run(200, 160, potential=lambda c, z: fam.potential_and_grad(z))
    last = _compare(out, fam, keys, 160, 40, find_heuristic_step_size=bool(heuristic), max_tree_depth=(5, 7))
    st, z, g, imm, sm = sim.state()
    np.testing.assert_array_equal(np.array(st[1].rng_key), last.rng_key)
    np.testing.assert_array_equal(np.array(st[1].adapt_rng_key), last.adapt_state.rng_key)
    np.testing.assert_array_equal(imm[1], last.adapt_state.inverse_mass_matrix)
