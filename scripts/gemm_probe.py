"""Quick timing of the GEMM regime at the BASELINE shapes (development aid; bench.py --config c3|c4 is the bench).

    python scripts/gemm_probe.py c3|c4 [passes]      # pass-bounded runs after a short warm-up + one potential-hook call
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from numpyro_b200 import _capi, engine as eng, random as b2random      # noqa: E402

F = np.float32


def data(which):
    rng = np.random.default_rng(33)
    if which == "c3":
        N, D, C = 100_000, 256, 16384
        X = (rng.standard_normal(size=(N, D), dtype=np.float32) / np.sqrt(D)).astype(F)
        X[:, 192:] = 0.0
        X[np.arange(N), 192 + rng.integers(0, 64, size=N)] = 1.0
        eta = np.clip(X @ (rng.normal(size=D) * 0.5), -10, 10)
        y = (rng.uniform(size=N) < 1 / (1 + np.exp(-eta))).astype(F)
        kw = dict(global_scale=_capi.SCALE_HALFCAUCHY, group_col_begin=192, group_col_end=256, tau_scale=1.0)
    else:
        N, D, C = 10_000, 1000, 1024
        X = rng.standard_normal(size=(N, D), dtype=np.float32)
        X -= X.mean(0)
        y = (2 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2] + 0.05 * rng.normal(size=N)).astype(F)
        kw = dict(likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY)
    return N, D, C, X, y, kw


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    N, D, C, X, y, kw = data(which)
    t0 = time.perf_counter()
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, max_tree_depth_warmup=8, max_tree_depth=8, **kw)
    torch.cuda.synchronize()
    print(f"{which}: N={N} D={D} C={C} regime={e.regime} info={e.gemm_info()} create {time.perf_counter() - t0:.2f} s", flush=True)
    z = (np.random.default_rng(1).normal(size=(C, e.D)) * 0.2).astype(F)
    e.potential_and_grad(z)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        e.potential_and_grad(z)
    ev1.record()
    torch.cuda.synchronize()
    print(f"potential hook (betas + gemm pass + finish): {ev0.elapsed_time(ev1) / 3:.3f} ms per call", flush=True)
    keys = b2random.split(b2random.PRNGKey(1), C)
    e.init(keys, 300)
    e.run(5, 5, fields=())                          # init + the first (tiny, divergent) trees
    lf = lambda: sum(int(s.total_leapfrogs) for s in e.state()[0])
    for rep in range(3):
        l0, p0 = lf(), e.pass_count
        c0 = e.debug_clocks().astype(np.float64)
        ev0.record()
        e.run(300, 300, fields=(), max_passes=passes)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        l1, p1 = lf(), e.pass_count
        c1 = e.debug_clocks().astype(np.float64)
        n = max(p1 - p0, 1)
        flops = 4.0 * N * D * (l1 - l0)
        print(f"run {rep}: {n} passes in {ms:.2f} ms = {ms / n:.3f} ms per pass; {l1 - l0} grad-evals "
              f"({(l1 - l0) / n:.0f} per pass) -> {(l1 - l0) / ms * 1e3:.0f} grad-evals/s, {flops / ms * 1e-9:.1f} algorithmic TFLOP/s; "
              f"CTA0: {(c1[0] - c0[0]) / n:.0f} cycles per pass in the kernel, {(c1[1] - c0[1]) / n:.1f} units, MMA thread waiting for the epilogue "
              f"{(c1[2] - c0[2]) / n:.0f} cycles, for tile copies {(c1[3] - c0[3]) / n:.0f}, for its own MMAs {(c1[4] - c0[4]) / n:.0f}", flush=True)
    st, _ = e.state()
    print("iterations reached:", min(s.i for s in st), max(s.i for s in st), "step sizes:", np.percentile([s.step_size for s in st], [0, 50, 100]))


if __name__ == "__main__":
    main()
