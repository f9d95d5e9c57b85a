"""Pins the oracle's HMC numerics with the reference's own known answers
(test/infer/test_hmc_util.py; line numbers cited per test)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import adapt as ad
from oracle import chain, detmath as dm, families, prng, tree

F = np.float32


# --- test_hmc_util.py:278-292 -------------------------------------------------------------
@pytest.mark.parametrize("num_steps, expected", [
    (18, [(0, 17)]),
    (50, [(0, 6), (7, 44), (45, 49)]),
    (100, [(0, 14), (15, 89), (90, 99)]),
    (150, [(0, 74), (75, 99), (100, 149)]),
    (200, [(0, 74), (75, 99), (100, 149), (150, 199)]),
    (280, [(0, 74), (75, 99), (100, 229), (230, 279)]),
    (1000, [(0, 74), (75, 99), (100, 149), (150, 249), (250, 449), (450, 949), (950, 999)]),
])
def test_build_adaptation_schedule(num_steps, expected):
    assert ad.build_adaptation_schedule(num_steps) == expected


# --- test_hmc_util.py:381-386 -------------------------------------------------------------
@pytest.mark.parametrize("leaf_idx, ckpt_idxs",
                         [(0, (1, 0)), (6, (3, 2)), (7, (0, 2)), (13, (2, 2)), (15, (0, 3))])
def test_leaf_idx_to_ckpt_idx(leaf_idx, ckpt_idxs):
    assert tree.leaf_idx_to_ckpt_idxs(leaf_idx) == ckpt_idxs


# --- test_hmc_util.py:389-403 -------------------------------------------------------------
@pytest.mark.parametrize("ckpt_idxs, expected_turning",
                         [((3, 2), False), ((3, 3), True), ((0, 0), False), ((0, 1), True), ((1, 3), True)])
def test_is_iterative_turning(ckpt_idxs, expected_turning):
    imm = np.ones(1, F)
    r_ckpts = np.array([[1.0], [2.0], [3.0], [-2.0]], F)
    r_sum_ckpts = np.array([[2.0], [4.0], [4.0], [-1.0]], F)
    got = tree.is_iterative_turning(imm, np.array([1.0], F), np.array([3.0], F), r_ckpts, r_sum_ckpts, *ckpt_idxs)
    assert got == expected_turning


# --- test_hmc_util.py:36-52 ---------------------------------------------------------------
def test_dual_averaging():
    s = ad.da_init(0.0)
    for _ in range(10):
        g = F(2.0) * (s.x_t + F(1.0))          # gradient of (x+1)^2
        s = ad.da_update(g, s, gamma=0.5)
    assert_allclose(s.x_avg, -1.0, atol=1e-3)


# --- test_hmc_util.py:55-89 (diagonal branch) ----------------------------------------------
@pytest.mark.parametrize("regularize", [True, False])
def test_welford_diag(regularize):
    rng = np.random.default_rng(0)
    cov = np.array([[1.0, 0.4, 0.0], [0.4, 2.0, -0.3], [0.0, -0.3, 0.5]])
    xs = rng.multivariate_normal(np.zeros(3), cov, size=2000).astype(F)
    s = ad.welford_init(3)
    for x in xs:
        s = ad.welford_update(x, s)
    est, sqrt_m, sqrt_inv = ad.welford_final(s, regularize)
    want = xs.astype(np.float64).var(axis=0, ddof=1)
    if regularize:
        want = want * (2000 / 2005) + 1e-3 * 5 / 2005
    assert_allclose(est, want, rtol=1e-4)
    assert_allclose(sqrt_inv * sqrt_inv, est, rtol=1e-5)
    assert_allclose(sqrt_m * sqrt_inv, 1.0, rtol=1e-6)


# --- test_hmc_util.py:121-229: leapfrog on analytic systems ---------------------------------
def _integrate(pot, eps, n, q, p, imm):
    u, g = pot(q)
    for _ in range(n):
        q, p, u, g = tree.leapfrog(pot, eps, imm, q, p, g)
    return q, p, u


def test_leapfrog_harmonic_oscillator():
    pot = lambda q: (F(0.5 * float(q[0]) ** 2), q.copy())
    imm = np.ones(1, F)
    q0, p0 = np.array([0.0], F), np.array([1.0], F)
    q, p, u = _integrate(pot, 0.01, 100, q0, p0, imm)
    assert_allclose(q[0], np.sin(1.0), atol=1e-4)
    assert_allclose(p[0], np.cos(1.0), atol=1e-4)
    assert_allclose(u + tree.kinetic_energy(imm, p), 0.5, atol=1e-5)
    qb, pb, _ = _integrate(pot, 0.01, 100, q, -p, imm)
    assert_allclose(qb, q0, atol=1e-4)


def test_leapfrog_circular_orbit():
    def pot(q):
        r = float(np.sqrt(np.sum(q.astype(np.float64) ** 2)))
        return F(-1.0 / r), (q.astype(np.float64) / r ** 3).astype(F)
    imm = np.ones(2, F)
    q0, p0 = np.array([1.0, 0.0], F), np.array([0.0, 1.0], F)
    q, p, u = _integrate(pot, 0.01, 628, q0, p0, imm)
    assert_allclose(q, [1.0, 0.0], atol=5e-3)
    assert_allclose(p, [0.0, 1.0], atol=5e-3)
    assert_allclose(u + tree.kinetic_energy(imm, p), -0.5, atol=1e-5)


def test_leapfrog_quartic():
    pot = lambda q: (F(0.25 * float(q[0]) ** 4), (q.astype(np.float64) ** 3).astype(F))
    imm = np.ones(1, F)
    q0, p0 = np.array([0.02], F), np.array([0.0], F)
    q, p, u = _integrate(pot, 0.1, 1810, q0, p0, imm)
    assert_allclose(q[0], -0.02, atol=1e-4)
    assert_allclose(p[0], 0.0, atol=1e-4)


# --- test_hmc_util.py:232-275 -------------------------------------------------------------
@pytest.mark.parametrize("init_step_size", [0.1, 10.0])
def test_find_reasonable_step_size(init_step_size, monkeypatch):
    pot = lambda q: (F(0.5 * float(q[0]) ** 2), q.copy())
    monkeypatch.setattr(prng, "normal", lambda k, n=None: np.ones(n, F))     # p_generator == 1.0
    z = np.array([0.0], F)
    u, g = pot(z)
    step = ad.find_reasonable_step_size(pot, np.ones(1, F), np.ones(1, F), z, u, g, init_step_size, prng.key(0))
    threshold = (-np.log(0.8) * 8) ** 0.25
    if init_step_size < threshold:
        assert step / 2 < threshold < step
    else:
        assert step * 2 > threshold > step


# --- test_hmc_util.py:304-378 -------------------------------------------------------------
def test_warmup_adapter_script():
    find = lambda step, imm, sm, z, pe, g, k: F(step * 4) if step < 1 else F(step / 4)
    num_steps = 150
    sched = ad.build_adaptation_schedule(num_steps)
    wa = ad.WarmupAdapter(num_steps, find)
    z = np.ones(3, F)
    s = wa.init(z, F(0), z, prng.key(0), 1.0)
    assert s.step_size == F(0.25) and s.window_idx == 0
    assert_allclose(s.inverse_mass_matrix, 1.0)
    step0 = s.step_size
    w = sched[0]
    for t in range(w[0], w[1] + 1):
        s = wa.update(t, 0.7 + 0.1 * t / (w[1] - w[0]), z, F(0), z, s)
    assert s.window_idx == 1 and s.step_size < step0
    assert_allclose(s.inverse_mass_matrix, 1.0)
    step1 = s.step_size
    w = sched[1]
    for t in range(w[0], w[1] + 1):
        s = wa.update(t, 0.8 + 0.1 * (t - w[0]) / (w[1] - w[0]), 2 * z, F(0), z, s)
    assert s.window_idx == 2 and s.step_size > step1
    assert_allclose(s.inverse_mass_matrix, 1e-3 * (5 / (w[1] + 1 - w[0] + 5)), atol=1e-7)
    imm2, step2 = s.inverse_mass_matrix.copy(), s.step_size
    w = sched[2]
    for t in range(w[0], w[1] + 1):
        s = wa.update(t, 0.8, (t * z).astype(F), F(0), z, s)
    assert s.window_idx == 3
    assert_allclose(s.step_size, step2 * 10, atol=1e-6)
    assert_allclose(s.inverse_mass_matrix, imm2)


# --- test_hmc_util.py:406-442 -------------------------------------------------------------
@pytest.mark.parametrize("step_size", [0.01, 1.0, 100.0])
def test_build_tree_invariants(step_size):
    pot = lambda q: (F(0.5 * float(q[0]) ** 2), q.copy())
    z, r = np.array([0.0], F), np.array([1.0], F)
    u, g = pot(z)
    t = tree.build_tree(pot, np.ones(1, F), F(step_size), prng.key(0), z, r, u, g, 10)
    assert t.num_proposals >= 2 ** (t.depth - 1)
    assert t.sum_accept <= t.num_proposals
    if t.depth < 10:
        assert t.turning or t.diverging
    if step_size > 10:
        assert t.diverging and t.num_proposals == 1
    if step_size < 0.1:
        assert t.num_proposals > 10


def test_gradients_by_finite_differences():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(40, 5))
    fams = [
        families.EightSchools([15., 10, 16, 11, 9, 11, 10, 18], [28., 8, -3, 7, -1, 1, 18, 12]),
        families.logistic_regression(X, rng.integers(0, 2, 40)),
        families.GLM(X * 0.3, rng.poisson(2.0, 40), likelihood="poisson"),
        families.horseshoe(X, rng.integers(0, 2, 40), "bernoulli"),
        families.horseshoe(X, rng.normal(size=40), "normal"),
        families.GLM(X, rng.integers(0, 2, 40), global_scale="exponential", group_cols=(2, 5), tau_scale=2.0),
        families.GLM(X, rng.integers(0, 2, 40), global_scale="halfcauchy", group_cols=(1, 4)),
    ]
    for fam in fams:
        z = rng.normal(size=fam.dim) * 0.5
        u, g = fam.potential64(z)
        h = 1e-6
        num = np.array([(fam.potential64(z + h * e)[0] - fam.potential64(z - h * e)[0]) / (2 * h)
                        for e in np.eye(fam.dim)])
        assert_allclose(g, num, rtol=1e-5, atol=1e-6, err_msg=fam.name)


def test_layout_sorted_and_trace_order():
    fam = families.horseshoe(np.zeros((3, 4)), np.zeros(3), "normal")
    assert [n for n, _ in fam.init_sites] == ["lambdas", "tau", "unscaled_betas", "prec_obs"]
    assert [(n, o, s) for n, o, s in fam.layout] == [("lambdas", 0, 4), ("prec_obs", 4, 1), ("tau", 5, 1),
                                                     ("unscaled_betas", 6, 4)]


def test_thinning_collection_indices():
    """fori_collect arithmetic (numpyro/util.py:368-403; test/test_util.py:18-68)."""
    fam = families.DiagGaussian(np.zeros(2), np.ones(2))
    k = chain.Kernel(fam.potential_and_grad)
    full, _ = chain.run_chain(k, fam, prng.key(3), 6, 11, thinning=1)
    k = chain.Kernel(fam.potential_and_grad)
    thin, _ = chain.run_chain(k, fam, prng.key(3), 6, 11, thinning=3)
    assert thin["z"].shape[0] == 3
    # start = lower + (upper-lower) % thinning = 6 + 2 ; slots hold iterations 10, 13, 16 -> sample idx 4, 7, 10
    np.testing.assert_array_equal(thin["z"], full["z"][[4, 7, 10]])
