"""Parity at BASELINE config 2's FULL size (N = 581012, D = 54):
  * a whole 8-chain NUTS run (init, adaptation, trees) bit-exact against the oracle driven by the engine's potential hook
    (one hook launch per oracle leapfrog);
  * the posterior of a longer run against an independent fp64 Laplace reference (Newton mode + curvature) within 4 MCSE;
  * leapfrog reversibility and energy conservation (the reference's test/infer/test_hmc_util.py:121-229 properties);
  * chain independence: the chains of an 8-chain run are bit-identical to the same chains inside a 16-chain run (two chain
    groups, rotating passes) -- test/infer/test_mcmc.py:868-914 at full size;
  * run-to-run determinism of a whole NUTS run."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng            # noqa: E402
from oracle import chain, diag, families, prng            # noqa: E402
from test_gpu_parity import FIELDS, assert_run_equal, device_potential      # noqa: E402

F = np.float32
N, D = 581012, 54


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(1)
    X = rng.standard_normal(size=(N, D), dtype=F)
    beta = (rng.normal(size=D) * 0.3).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
    return torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()


def test_whole_run_bit_exact_at_full_size(data):
    """8 chains x 24 transitions (12 warm-up incl. the first adaptation window, 12 samples) at N = 581012: tree depths, PRNG
    streams, adaptation and every collected field bit-identical to the oracle (test_mcmc.py:868-914 style, full size)."""
    X, y = data
    kw = dict(max_tree_depth_warmup=6, max_tree_depth=6)
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=8, X=X, y=y, **kw)
    hook = eng.Engine(family=_capi.FAMILY_GLM, num_chains=8, X=X, y=y, **kw)      # a second handle serves the oracle's gradients
    assert e.regime == _capi.REGIME_STREAM
    keys = prng.split(prng.key(21), 8)
    e.init(keys, 12)
    out = e.run(24, 12, fields=FIELDS)

    class Fam:                                  # the oracle only needs the latent layout of the plain GLM
        init_sites = [("coefs", D)]
        layout = [("coefs", 0, D)]
    total = 0
    for c in range(8):
        kern = chain.Kernel(device_potential(hook, c), max_tree_depth=(6, 6))
        res, st = chain.run_chain(kern, Fam, keys[c], 12, 12, fields=FIELDS)
        assert_run_equal(out, res, c)
        total += int(np.sum(res["num_steps"]))
    assert total >= 8 * 12
    assert int(e.debug_clocks()[13]) == 0        # positions published ahead of the tick (Tick::peek_next) == the state machine's
    e.close(); hook.close()


def test_posterior_against_independent_laplace_reference(data):
    """Posterior means within 4 Monte-Carlo standard errors of an independent fp64 reference, split R-hat < 1.01.  With
    581012 rows the posterior is Gaussian to O(1/N): mean = mode (Newton's method in fp64), covariance = H^-1."""
    X, y = data
    X64, y64 = X.cpu().numpy().astype(np.float64), y.cpu().numpy().astype(np.float64)
    b = np.zeros(D)
    for _ in range(10):
        p = 1.0 / (1.0 + np.exp(-(X64 @ b)))
        H = (X64 * (p * (1 - p))[:, None]).T @ X64 + np.eye(D)
        b = b - np.linalg.solve(H, X64.T @ (p - y64) + b)
    cov = np.linalg.inv(H)
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=8, X=X, y=y)
    e.init(prng.split(prng.key(22), 8), 400)
    out = e.run(900, 400, fields=("z", "diverging"))
    z = out["z"].cpu().numpy().astype(np.float64)                   # [8, 500, 54]
    assert int(out["diverging"].sum().item()) == 0
    dbg = e.debug_clocks()
    assert int(dbg[11]) > 10 * int(dbg[12]) and int(dbg[13]) == 0   # early publishes dominate post warm-up, and never disagree
    ess = diag.effective_sample_size(z)
    sd = z.std(axis=(0, 1))
    mcse = sd / np.sqrt(ess)
    assert np.all(diag.split_gelman_rubin(z) < 1.01)
    assert np.all(np.abs(z.mean(axis=(0, 1)) - b) < 4 * mcse + 1e-6), np.max(np.abs(z.mean(axis=(0, 1)) - b) / mcse)
    np.testing.assert_allclose(sd, np.sqrt(np.diag(cov)), rtol=0.1)            # posterior scale = Laplace scale
    e.close()


def test_leapfrog_reversible_and_energy_conserving_at_full_size(data):
    X, y = data
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=8, X=X, y=y)
    assert e.regime == _capi.REGIME_STREAM
    rng = np.random.default_rng(3)
    z0 = (rng.normal(size=(8, D)) * 0.05).astype(F)
    imm = np.full((8, D), 2e-6, F)                       # ~ posterior variance scale of this dataset
    r0 = (rng.normal(size=(8, D)) / np.sqrt(imm)).astype(F)
    eps = np.full(8, 0.05, F)
    U0, _ = e.potential_and_grad(z0)
    z1, r1, U1, _ = e.leapfrog(eps, imm, z0, r0, 12)
    kin = lambda r: 0.5 * (imm * r.astype(np.float64) ** 2).sum(1)
    E0 = U0.cpu().numpy().astype(np.float64) + kin(r0)
    E1 = U1.cpu().numpy().astype(np.float64) + kin(r1.cpu().numpy())
    np.testing.assert_allclose(E1, E0, rtol=2e-5)        # energy drift of a stable trajectory (fp32 potential)
    zb, rb, Ub, _ = e.leapfrog(eps, imm, z1.cpu().numpy(), -r1.cpu().numpy(), 12)
    np.testing.assert_allclose(zb.cpu().numpy(), z0, rtol=1e-4, atol=1e-5)          # reversibility <= 1e-4
    np.testing.assert_allclose(-rb.cpu().numpy(), r0, rtol=1e-4, atol=1e-4 * np.abs(r0).max())
    np.testing.assert_allclose(Ub.cpu().numpy(), U0.cpu().numpy(), rtol=1e-6)
    e.close()


def test_chain_independence_and_determinism_at_full_size(data):
    X, y = data
    keys = prng.split(prng.key(5), 16)
    fields = ("z", "num_steps", "potential_energy")

    def run(C, ks):
        e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, max_tree_depth_warmup=6, max_tree_depth=6)
        e.init(ks, 12)
        out = e.run(18, 12, fields=fields)
        st, _ = e.state()
        e.close()
        return out, [int(s.total_leapfrogs) for s in st]
    o16, l16 = run(16, keys)
    o8a, l8a = run(8, keys[:8])
    o8b, l8b = run(8, keys[8:])
    o8a2, _ = run(8, keys[:8])
    for f in fields:
        assert torch.equal(o16[f][:8], o8a[f]) and torch.equal(o16[f][8:], o8b[f]), f
        assert torch.equal(o8a[f], o8a2[f]), f
    assert l16 == l8a + l8b and min(l16) >= 18
    assert torch.isfinite(o16["z"]).all()
