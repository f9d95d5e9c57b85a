"""Parity at BASELINE config 2's FULL size (N = 581012, D = 54) through size-independent properties -- the oracle cannot
run whole chains at this size in test time:
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
from oracle import prng                                   # noqa: E402

F = np.float32
N, D = 581012, 54


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(1)
    X = rng.standard_normal(size=(N, D), dtype=F)
    beta = (rng.normal(size=D) * 0.3).astype(F)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
    return torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()


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
