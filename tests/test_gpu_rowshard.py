"""Row-sharded handles (BASELINE config 5): every rank sweeps its own rows, the per-chain likelihood
sums are all-reduced inside the kernel (peer stores into the ranks' mailboxes, added in rank order) and all ranks
must see bit-identical (U, grad) and produce bit-identical chains.

On a one-GPU box the two ranks are two threads that share the device (each handle takes half of the SMs through the
B200NUTS_GRID testing aid); with >= 2 GPUs the same test also runs one rank per device, and a torchrun variant
exercises the CUDA-IPC path (ranks = processes)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from numpyro_b200 import _capi, engine as eng          # noqa: E402
from oracle import chain, families, prng                # noqa: E402

F = np.float32
FIELDS = ("z", "num_steps", "accept_prob", "potential_energy", "diverging", "step_size")


def _data(N, D, seed, lik="bernoulli"):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D)).astype(F)
    beta = rng.normal(size=D) * 0.3
    if lik == "poisson":
        X = (X * 0.3).astype(F)
        y = rng.poisson(np.exp(np.clip(X @ beta, -3, 3))).astype(F)
    else:
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ beta))).astype(F)
    return X, y


class Ranks:
    """W row-sharded handles driven by W threads of this process."""

    def __init__(self, X, y, C, cuts, devices, regime=_capi.REGIME_STREAM, **cfg):
        self.W = len(cuts) - 1
        self.streams, self.engines = [], []
        for r in range(self.W):
            dev = torch.device("cuda", devices[r])
            torch.cuda.set_device(dev)
            e = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=C, X=X[cuts[r]:cuts[r + 1]], y=y[cuts[r]:cuts[r + 1]],
                           regime=regime, shard_rank=r, shard_count=self.W, n_rows_global=X.shape[0], **cfg)
            self.engines.append(e)
            self.streams.append(torch.cuda.Stream(device=dev))
        blobs = [e.shard_blob() for e in self.engines]
        for e in self.engines:
            e.connect_shards(blobs)

    def each(self, fn):
        out, err = [None] * self.W, []

        def work(r):
            try:
                torch.cuda.set_device(self.engines[r].device)
                with torch.cuda.stream(self.streams[r]):
                    out[r] = fn(self.engines[r])
                    self.streams[r].synchronize()
            except Exception as ex:                      # noqa: BLE001
                err.append(ex)
        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.W)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            same_device = len({e.device for e in self.engines}) == 1
            if same_device and any("never arrived" in str(x) or "watchdog" in str(x) or "timed out" in str(x) for x in err):
                # two persistent grids on ONE device only work when the hardware runs them side by side
                pytest.skip("the two half-grid handles were not co-resident on this device")
            raise err[0]
        return out

    def close(self):
        for e in self.engines:
            e.close()


def _layouts():
    n = torch.cuda.device_count()
    out = [pytest.param((0, 0), id="two-ranks-one-device")]
    if n >= 2:
        out.append(pytest.param((0, 1), id="two-devices"))
    return out


@pytest.fixture
def half_grid(monkeypatch):
    monkeypatch.setenv("B200NUTS_GRID", str(torch.cuda.get_device_properties(0).multi_processor_count // 2))
    monkeypatch.setenv("B200NUTS_WATCHDOG_S", "60")


@pytest.mark.parametrize("devices", _layouts())
@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_sharded_potential_identical_on_all_ranks_and_equal_to_unsharded(devices, lik, half_grid):
    N, D, C = 30011, 54, 5
    X, y = _data(N, D, 3, lik)
    kw = dict(likelihood=_capi.LIK_POISSON_LOG) if lik == "poisson" else {}
    rk = Ranks(X, y, C, [0, 17003, N], devices, **kw)
    rng = np.random.default_rng(0)
    z = (rng.normal(size=(C, D)) * 0.3).astype(F)
    res = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(z)))
    (U0, g0), (U1, g1) = res
    assert np.array_equal(U0, U1) and np.array_equal(g0, g1), "ranks disagree: replicated chains would diverge"
    fam = families.GLM(X, y, likelihood="poisson") if lik == "poisson" else families.logistic_regression(X, y)
    for c in range(C):
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U0[c], u64, rtol=1e-5)
        np.testing.assert_allclose(g0[c], g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max())
    # same bits again on a second launch (the exchange tags carry the launch epoch)
    res2 = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(z)))
    assert np.array_equal(res2[0][0], U0) and np.array_equal(res2[1][1], g0)
    rk.close()


@pytest.mark.parametrize("C", [3, 12])                    # 12 chains = two chain groups with rotating passes
@pytest.mark.parametrize("devices", _layouts())
def test_sharded_run_bit_identical_across_ranks_and_bit_exact_against_oracle(devices, C, half_grid):
    N, D = 9000, 7
    X, y = _data(N, D, 4)
    rk = Ranks(X, y, C, [0, 4000, N], devices, max_tree_depth_warmup=5, max_tree_depth=5)
    keys = prng.split(prng.key(7), C)
    rk.each(lambda e: e.init(keys, 40))
    outs = rk.each(lambda e: {k: v.cpu().numpy() for k, v in e.run(60, 40, fields=FIELDS).items()})
    for f in FIELDS:
        assert np.array_equal(outs[0][f], outs[1][f]), f
    fam = families.logistic_regression(X, y)

    def device_potential(c):
        def pot(zc):
            zz = np.zeros((C, D), F)
            zz[c] = zc
            U, g = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(zz)))[0]
            return F(U[c]), g[c]
        return pot
    cc = C - 2
    kern = chain.Kernel(device_potential(cc), max_tree_depth=(5, 5))
    res, _ = chain.run_chain(kern, fam, keys[cc], 40, 20, fields=FIELDS)
    for f in FIELDS:
        assert np.array_equal(outs[0][f][cc], res[f]), f
    rk.close()


# ---- the many-chain GEMM regime with the rows split over ranks (BASELINE config 5: 64 chains x 128 columns): the all-reduce
#      of the per-chain sums happens in the tick kernel that follows every GEMM pass (gemm_shard_allreduce)
@pytest.mark.parametrize("devices", _layouts())
@pytest.mark.parametrize("lik", ["bernoulli", "poisson"])
def test_gemm_sharded_potential_identical_on_all_ranks_and_equal_to_fp64(devices, lik, half_grid):
    N, D, C = 20011, 100, 40
    X, y = _data(N, D, 13, lik)
    kw = dict(likelihood=_capi.LIK_POISSON_LOG) if lik == "poisson" else {}
    rk = Ranks(X, y, C, [0, 11003, N], devices, regime=_capi.REGIME_GEMM, **kw)
    assert all(e.regime == _capi.REGIME_GEMM for e in rk.engines)
    rng = np.random.default_rng(0)
    z = (rng.normal(size=(C, D)) * 0.2).astype(F)
    (U0, g0), (U1, g1) = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(z)))
    assert np.array_equal(U0, U1) and np.array_equal(g0, g1), "ranks disagree: replicated chains would diverge"
    fam = families.GLM(X, y, likelihood="poisson") if lik == "poisson" else families.logistic_regression(X, y)
    for c in (0, 17, C - 1):
        u64, g64 = fam.potential64(z[c].astype(np.float64))
        np.testing.assert_allclose(U0[c], u64, rtol=1e-5)
        np.testing.assert_allclose(g0[c], g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max())
    res2 = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(z)))      # tags carry the launch number
    assert np.array_equal(res2[0][0], U0) and np.array_equal(res2[1][1], g0)
    rk.close()


@pytest.mark.parametrize("devices", _layouts())
def test_gemm_sharded_run_bit_identical_across_ranks_and_bit_exact_against_oracle(devices, half_grid):
    N, D, C = 9000, 70, 24
    X, y = _data(N, D, 14)
    rk = Ranks(X, y, C, [0, 5000, N], devices, regime=_capi.REGIME_GEMM, max_tree_depth_warmup=4, max_tree_depth=4)
    keys = prng.split(prng.key(9), C)
    rk.each(lambda e: e.init(keys, 12))
    # two launches (12 warm-up transitions collected away, then 8 samples; the second is pass-bounded and resumed)
    rk.each(lambda e: e.run(12, 12, fields=()))
    outs = rk.each(lambda e: e.run(20, 12, fields=FIELDS, max_passes=25))
    outs = rk.each(lambda e: {k: v.cpu().numpy() for k, v in e.run(20, 12, fields=FIELDS, out=outs[rk.engines.index(e)]).items()})
    for f in FIELDS:
        assert np.array_equal(outs[0][f], outs[1][f]), f
    fam = families.logistic_regression(X, y)

    def device_potential(c):
        def pot(zc):
            zz = np.zeros((C, D), F)
            zz[c] = zc
            U, g = rk.each(lambda e: tuple(t.cpu().numpy() for t in e.potential_and_grad(zz)))[0]
            return F(U[c]), g[c]
        return pot
    cc = 5
    kern = chain.Kernel(device_potential(cc), max_tree_depth=(4, 4))
    res, _ = chain.run_chain(kern, fam, keys[cc], 12, 8, fields=FIELDS)
    for f in FIELDS:
        assert np.array_equal(outs[0][f][cc], res[f]), f
    rk.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (ranks = processes, CUDA IPC)")
@pytest.mark.parametrize("regime", ["stream", "gemm"])
def test_sharded_processes_over_cuda_ipc(regime):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "rowshard_worker.py"), regime]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "ROWSHARD_OK" in p.stdout


@pytest.mark.parametrize("devices", _layouts())
def test_mcmc_row_shards_matches_engine_level_ranks(devices, half_grid):
    """Public API: MCMC(NUTS(LogisticRegression()), row_shards=2) == the engine-level ranks over the same row cuts."""
    if devices == (0, 1):
        pytest.skip("the API picks cuda:0..G-1 itself; covered by the one-device layout when only the mailboxes differ")
    from numpyro_b200 import families as model_families, random as b2random
    from numpyro_b200.infer import MCMC, NUTS
    N, D, C = 20000, 12, 4
    X, y = _data(N, D, 9)
    mcmc = MCMC(NUTS(model_families.LogisticRegression(), max_tree_depth=5), num_warmup=30, num_samples=20, num_chains=C,
                chain_method="vectorized", progress_bar=False, row_shards=2)
    mcmc.run(b2random.PRNGKey(4), X, y, extra_fields=("num_steps",))
    got = mcmc.get_samples(group_by_chain=True)["coefs"]
    steps = mcmc.get_extra_fields(group_by_chain=True)["num_steps"]
    assert got.shape == (C, 20, D) and np.isfinite(got).all()
    for s in mcmc._shards:
        s.engine.close()
    rk = Ranks(X, y, C, [0, N // 2, N], (0, 0), max_tree_depth_warmup=5, max_tree_depth=5)
    keys = prng.split(prng.key(4), C)
    rk.each(lambda e: e.init(keys, 30))
    outs = rk.each(lambda e: {k: v.cpu().numpy() for k, v in e.run(50, 30, fields=("z", "num_steps")).items()})
    rk.close()
    assert np.array_equal(outs[0]["z"], got) and np.array_equal(outs[0]["num_steps"], steps)
    # and the posterior is the logistic-regression posterior: coefficient means near the maximum-likelihood direction
    assert np.abs(got.mean((0, 1))).max() < 2.0
