#!/usr/bin/env python
"""Benchmarks of the NUTS hot path: gradient evaluations per second (and min-ESS per second).

    python bench.py --gpus N --steps K --warmup W [--config c2]     # this repo (CUDA engine)
    python bench.py --impl reference --gpus N ... [--config c2]     # CPU arm: the oracle port on the host cores

Configs (BASELINE.json `configs`); the default and the headline is c2:
  c1  eight schools non-centred, 4 chains, 1000/1000 (README example): transitions/s + posterior table
  c2  covtype-shaped Bayesian logistic regression NUTS, N = 581012, D = 54 fp32, 8 chains per GPU  -- streaming regime, HBM roofline
  c3  hierarchical GLM, N = 100k, D = 256, 16384 chains per GPU                                    -- tcgen05 GEMM regime, tensor roofline
  c4  horseshoe regression, N = 10k, D = 1000, 1024 chains per GPU                                 -- tcgen05 GEMM regime
  c5  row-sharded logistic regression, 25M rows x 128 columns PER GPU (200M over 8), 64 replicated chains, the per-chain
      (log-density, gradient) sums all-reduced over the ranks after every pass                     -- GEMM regime + peer mailboxes

c2.  A *step* is TRANSITIONS_PER_STEP NUTS transitions of every chain (fixed samples per chain, what a user of
``MCMC.run`` experiences), post warm-up; adaptation runs before the timed region as setup.  A GPU needs as many sweeps of X as
its slowest chain needs leapfrogs, so `value` includes that imbalance; the pass-bounded throughput (every GPU runs the same
number of sweeps, chains pause mid-tree) is reported next to it as `pass_bounded`.  `value` is measured with the dataset
resident in HBM; `e2e` runs the public ``MCMC`` API from pinned host buffers (H2D of X and y, engine creation, init,
warm-up, sampling, D2H of the samples) inside the timed region.
c3 / c4.  A step is PASSES_PER_STEP gradient passes over all chains (one GEMM-regime launch = gemm pass + tick + schedule per
pass inside a device-side WHILE graph).
Multi-GPU (torchrun, one rank per GPU): chains shard across ranks with no data-path collective (weak scaling, dataset
replicated); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F = np.float32
WORKLOADS = {
    "c1": "configs[0]: eight_schools non-centered NUTS, 4 chains, 1000 warmup / 1000 samples",
    "c2": "configs[1]: covtype-shaped Bayesian logistic regression NUTS (N=581012, D=54 fp32, synthetic, 8 chains per GPU, max_tree_depth=10)",
    "c3": "configs[2]: vectorized-chain hierarchical GLM NUTS (N=100000, D=256 with a 64-column group block sharing a HalfCauchy scale, Bernoulli-logit, synthetic, 16384 chains per GPU)",
    "c4": "configs[3]: horseshoe regression NUTS (examples/horseshoe_regression.py recipe scaled to N=10000, D=1000, Normal likelihood, synthetic, 1024 chains per GPU)",
    "c5": "configs[4]: row-sharded logistic regression NUTS (N=25M rows per GPU = 200M over 8 GPUs, D=128 fp32, generated on the device, 64 chains replicated on every GPU, all-reduce of grad+logp per leapfrog)",
}
C2_ROWS, C2_COLS, C2_CHAINS = 581012, 54, 8
C2_BYTES_PER_PASS = C2_ROWS * C2_COLS * 4 + C2_ROWS * 4          # one sweep of X and y serves every chain
C2_TRANSITIONS_PER_STEP = 400
C2_PASSES_PER_STEP = 3000           # pass-bounded secondary measurement
C2_ADAPT_ITERS = 600
C5 = dict(rows_per_gpu=25_000_000, D=128, C=64, warm_passes=12, passes_per_step=4, num_warmup=100)
GEMM = {"c3": dict(N=100_000, D=256, C=16384, warm_passes=120, passes_per_step=12),
        "c4": dict(N=10_000, D=1000, C=1024, warm_passes=300, passes_per_step=60)}


# ------------------------------------------------------------------------------------------ data
def make_data(config):
    if config == "c2":
        rng = np.random.default_rng(1)
        X = rng.standard_normal(size=(C2_ROWS, C2_COLS), dtype=np.float32)
        X = (X - X.mean(0)) / X.std(0)                        # column-standardised as examples/covtype.py:48
        beta = (rng.normal(size=C2_COLS) * 0.3).astype(F)
        y = (rng.uniform(size=C2_ROWS) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(F)
        return np.ascontiguousarray(X, F), y
    if config == "c3":
        g = GEMM["c3"]
        rng = np.random.default_rng(33)
        X = (rng.standard_normal(size=(g["N"], g["D"]), dtype=np.float32) / np.sqrt(g["D"])).astype(F)
        X[:, 192:] = 0.0
        X[np.arange(g["N"]), 192 + rng.integers(0, 64, size=g["N"])] = 1.0          # one-hot group block: random intercepts
        eta = np.clip(X @ (rng.normal(size=g["D"]) * 0.5), -10, 10)
        return X, (rng.uniform(size=g["N"]) < 1 / (1 + np.exp(-eta))).astype(F)
    if config == "c4":
        g = GEMM["c4"]
        rng = np.random.default_rng(0)                        # examples/horseshoe_regression.py:105-125
        X = rng.standard_normal(size=(g["N"], g["D"]), dtype=np.float32)
        X -= X.mean(0)
        return X, (2 * X[:, 0] - X[:, 1] + 0.5 * X[:, 2] + 0.05 * rng.normal(size=g["N"])).astype(F)
    if config == "c5":                                            # CPU legs only: a 200k-row sample of one GPU's shard
        rng = np.random.default_rng(5)
        X = rng.standard_normal(size=(200_000, C5["D"]), dtype=np.float32)
        beta = (rng.normal(size=C5["D"]) * 0.1).astype(F)
        return X, (rng.uniform(size=X.shape[0]) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
    raise ValueError(config)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops_sustained": float(d["bf16_tflops_sustained"]),
                "bf16_tflops": float(d["bf16_tflops"]), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0,
            "source": "B200_PROFILING.md fallback 6.65 TB/s / 1.59 PFLOP/s burst, ~1.4 sustained (of fallback)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
_CPU = {}          # per-process cache: dataset (inherited through fork, never pickled) and the family built from it


def _cpu_pool_init():
    """One BLAS thread per worker: `cores` workers x multithreaded BLAS oversubscribed the host 4x in round 1."""
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(limits=1)
    except Exception:       # noqa: BLE001
        pass


def _cpu_family(config):
    from oracle import families
    if "fam" not in _CPU:
        X, y = _CPU["X"], _CPU["y"]
        if config == "c2":
            fam = families.logistic_regression(X, y)
            _CPU["pot"] = fam.potential_and_grad_f32          # fp32 BLAS matvecs: the arithmetic a CPU run of the reference performs
        elif config == "c5":
            fam = families.logistic_regression(X, y)
            _CPU["pot"] = fam.potential_and_grad_f32
        elif config == "c3":
            fam = families.GLM(X, y, global_scale="halfcauchy", group_cols=(192, 256), tau_scale=1.0)
            _CPU["pot"] = fam.potential_and_grad
        else:
            fam = families.horseshoe(X, y, "normal")
            _CPU["pot"] = fam.potential_and_grad
        _CPU["fam"] = fam
    return _CPU["pot"]


def _cpu_worker(args):
    """One oracle NUTS transition (tree depth capped so it finishes) on one host core."""
    config, z0, step, imm, key, depth = args
    from oracle import prng
    from oracle.tree import build_tree
    pot = _cpu_family(config)
    u, g = pot(z0)
    r = (1.0 / np.sqrt(imm) * prng.normal(key, z0.shape[0])).astype(F)
    t0 = time.perf_counter()
    tree = build_tree(pot, imm, F(step), key, z0, r, u, g, depth)
    return tree.num_proposals, time.perf_counter() - t0


def cpu_start(config, X, y, chains=8):
    """Deterministic start for the CPU legs (the same one in `cpu_baseline` and in --impl reference): for c2 a posterior
    draw under the Laplace approximation (mode by Newton's method in fp64, M^-1 = diag(H^-1)) with a typical adapted step
    size; for c3 / c4 a point near the origin with a small step."""
    rng = np.random.default_rng(7)
    if config == "c2":
        X64, y64 = X.astype(np.float64), y.astype(np.float64)
        b = np.zeros(X.shape[1])
        for _ in range(8):
            p = 1.0 / (1.0 + np.exp(-(X64 @ b)))
            H = (X64 * (p * (1 - p))[:, None]).T @ X64 + np.eye(X.shape[1])
            b = b - np.linalg.solve(H, X64.T @ (p - y64) + b)
        var = np.diag(np.linalg.inv(H))
        z = (b[None] + rng.normal(size=(chains, X.shape[1])) * np.sqrt(var)[None]).astype(F)
        return z, np.full(chains, 0.33, F), np.tile(var.astype(F), (chains, 1))
    D = X.shape[1] + 1 if config == "c3" else X.shape[1] if config == "c5" else 2 * X.shape[1] + 2
    z = (rng.normal(size=(chains, D)) * 0.05).astype(F)
    return z, np.full(chains, 0.01, F), np.full((chains, D), 1.0, F)


def cpu_rounds(config, X, y, cores, depth, rounds):
    """`rounds` timed rounds, each `cores` single-threaded processes x 1 oracle NUTS transition (tree depth <= `depth`)
    from the same start; returns the per-round grad-evals/s (wall clock of the round)."""
    import multiprocessing as mp
    from oracle import prng
    z, step, imm = cpu_start(config, X, y)
    _CPU.clear()
    _CPU["X"], _CPU["y"] = X, y
    keys = prng.split(prng.key(123), cores * (rounds + 1))
    job = lambda i, d: (config, z[i % z.shape[0]], float(step[i % len(step)]), imm[i % imm.shape[0]], keys[i], d)
    rates, leaps = [], []
    with mp.get_context("fork").Pool(cores, initializer=_cpu_pool_init) as pool:
        pool.map(_cpu_worker, [job(i, 1) for i in range(cores)])                     # start-up, fp32 copies, BLAS warm-up
        for r in range(rounds):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [job((r + 1) * cores + i, depth) for i in range(cores)])
            wall = time.perf_counter() - t0
            leaps.append(sum(v[0] for v in res))
            rates.append(leaps[-1] / wall)
    return rates, leaps


def cpu_baseline(config, X, y, depth=5, rounds=3):
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    rates, leaps = cpu_rounds(config, X, y, cores, depth, rounds)
    med = float(np.median(rates))
    pot = "fp32 BLAS potential" if config == "c2" else "fp64 NumPy potential rounded once"
    return {"value": med, "unit": "grad-evals/s", "cores": cores, "kind": "port",
            "rounds": [round(v, 2) for v in rates], "spread": (max(rates) - min(rates)) / med,
            "sample": f"{rounds} rounds x {cores} single-threaded processes x 1 oracle NUTS transition from the same Laplace / "
                      f"fixed start, tree depth capped at {depth} ({sum(leaps)} leapfrogs in total), {pot}, one BLAS thread per "
                      f"process; median of the rounds; {time.perf_counter() - t0:.1f} s wall incl. start-up"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    config = args.config
    if config == "c1":
        return run_c1_reference(args)
    X, y = make_data(config)
    cores = os.cpu_count() or 1
    rates, leaps = cpu_rounds(config, X, y, cores, depth=5 if config == "c2" else 3, rounds=max(args.steps, 3))
    value = float(np.median(rates))
    sample = (f"{len(rates)} steps, each {cores} single-threaded processes x 1 oracle NUTS transition from the same start "
              f"(tree depth capped, {sum(leaps)} leapfrogs in total); median; spread {(max(rates) - min(rates)) / value:.2f}")
    if config == "c5":
        scale = X.shape[0] / float(C5["rows_per_gpu"] * max(args.gpus, 1))
        sample += f"; run on {X.shape[0]} rows and scaled linearly to the {C5['rows_per_gpu'] * max(args.gpus, 1)} rows of the config (x {scale:.3g})"
        rates = [r * scale for r in rates]
        value *= scale
    line = {"impl": "reference", "metric": "grad_evals_per_sec", "value": value, "unit": "grad-evals/s",
            "n_gpus": args.gpus, "steps": len(rates), "warmup": 1, "ms_per_step": 1e3 * float(np.median(leaps)) / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[config], "arm": "restated reference (oracle port; jax/numpyro are not installable here)"},
            "cpu_baseline": {"value": value, "unit": "grad-evals/s", "cores": cores, "kind": "port", "sample": sample,
                             "rounds": [round(v, 2) for v in rates]},
            "e2e": {"value": value, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ shared GPU plumbing
class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def gather_obj(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed(torch, fn):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)


# ------------------------------------------------------------------------------------------ c2
def run_c2(args):
    import torch
    from numpyro_b200 import _capi, diagnostics, engine as eng, families, random as b2random
    from numpyro_b200.infer import MCMC, NUTS
    d = Dist()
    world, rank, dev = d.world, d.rank, d.dev
    C = C2_CHAINS
    X, y = make_data("c2")
    Xp, yp = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    keys = b2random.split(b2random.PRNGKey(1), C * world)[rank * C:(rank + 1) * C]

    # ---- setup (untimed): dataset to HBM, chain init, warm-up adaptation
    e = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=C, X=Xp, y=yp)
    assert e.regime == _capi.REGIME_STREAM
    e.init(keys, C2_ADAPT_ITERS)
    e.run(C2_ADAPT_ITERS, C2_ADAPT_ITERS, fields=())
    fields = ("z", "num_steps", "diverging")
    T, W = C2_TRANSITIONS_PER_STEP, max(args.warmup, 3)
    pos = [C2_ADAPT_ITERS]

    def step():
        out = e.run(pos[0] + T, pos[0], fields=fields)
        pos[0] += T
        return out

    for _ in range(W):
        step()
    clocks = ClockSampler(d.local)
    if rank == 0:
        clocks.start()
    d.barrier()
    l0, p0 = e.launch_count, e.pass_count
    outs = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        outs.append(step())
    ev1.record()
    d.barrier()
    ms = ev0.elapsed_time(ev1)
    launches, passes = e.launch_count - l0, e.pass_count - p0
    cyc_per_pass = float(e.debug_clocks()[5]) / max(passes / max(args.steps, 1), 1.0)      # clock64 on CTA 0, last launch
    z = torch.cat([o["z"] for o in outs], dim=1)                                          # [C, K*T, D]
    leap = int(sum(int(o["num_steps"].sum().item()) for o in outs))
    per_chain = torch.stack([o["num_steps"].sum(dim=1) for o in outs]).sum(dim=0).tolist()
    diverging = int(sum(int(o["diverging"].sum().item()) for o in outs))
    st, vec = e.state()

    # ---- secondary: pass-bounded steps (every GPU runs the same number of sweeps; chains pause mid-tree and resume)
    lf = lambda: sum(int(s.total_leapfrogs) for s in e.state()[0])
    window = (pos[0], pos[0] + 10 * C2_PASSES_PER_STEP)
    pb_out = e.run(window[1], window[0], fields=("num_steps",), max_passes=C2_PASSES_PER_STEP)
    d.barrier()
    lf0, pp0 = lf(), e.pass_count
    pb_ms = timed(torch, lambda: [e.run(window[1], window[0], fields=("num_steps",), max_passes=C2_PASSES_PER_STEP, out=pb_out) for _ in range(2)])
    d.barrier()
    pb_leap, pb_passes = lf() - lf0, e.pass_count - pp0

    # ---- end to end through the public API (pinned host inputs -> samples on the host), numpyro's covtype proportions
    e2e_warm, e2e_samples = 500, 500
    model = families.LogisticRegression()
    e2e_leap, e2e_ms, d2h, items = 0, 0.0, 0, {}
    for it in range(2):                                               # first iteration is a warm-up
        d.barrier()
        t0 = time.perf_counter()
        mcmc = MCMC(NUTS(model), num_warmup=e2e_warm, num_samples=e2e_samples, num_chains=C, chain_method="vectorized", progress_bar=False)
        mcmc.run(keys, Xp, yp, extra_fields=("num_steps",))
        samples = mcmc.get_samples()
        d.barrier()
        if it > 0:
            e2e_ms = (time.perf_counter() - t0) * 1e3
            e2e_leap = mcmc.total_grad_evals
            d2h = sum(v.nbytes for v in samples.values()) + mcmc.get_extra_fields()["num_steps"].nbytes
            items = dict(mcmc.timings)
        for s in mcmc._shards:
            s.engine.close()

    # ---- secondary measurement: the same dataset with 32 chains per GPU (chain groups rotate, ticks hide behind sweeps)
    many = None
    if not args.no_many_chains:
        CM = 32
        keys_m = b2random.split(b2random.PRNGKey(2), CM * world)[rank * CM:(rank + 1) * CM]
        em = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=CM, X=e.X, y=e.y)
        em.init(keys_m, 300)
        em.run(300, 300, fields=())
        em.run(350, 300, fields=("num_steps",))
        d.barrier()
        pm0 = em.pass_count
        om = {}
        mms = timed(torch, lambda: om.update(em.run(550, 350, fields=("num_steps",))))
        d.barrier()
        mpasses, mleap = em.pass_count - pm0, int(om["num_steps"].sum().item())
        many = {"chains_per_gpu": CM, "transitions_per_chain": 200, "grad_evals_per_sec_this_gpu": mleap / (mms * 1e-3),
                "us_per_pass": mms * 1e3 / max(mpasses, 1), "grad_evals_per_pass": mleap / max(mpasses, 1),
                "roofline_frac": mpasses * C2_BYTES_PER_PASS / (mms * 1e-3) / 1e9 / measured_peaks()["hbm_gbs"]}
        em.close()

    # ---- reduce over ranks
    mx = d.reduce([ms, e2e_ms, pb_ms], "MAX")
    sm = d.reduce([float(leap), float(e2e_leap), float(diverging), float(pb_leap), float(passes)], "SUM")
    per_rank = d.gather_obj({"rank": rank, "ms": ms, "passes": int(passes), "grad_evals": leap, "grad_evals_per_chain": per_chain,
                             "step_size": [round(float(s_.step_size), 5) for s_ in st], "pass_bounded_ms": pb_ms})
    if world > 1:
        zs = [torch.empty_like(z) for _ in range(world)]
        d.dist.all_gather(zs, z)                                       # final gather for the diagnostics only
        z = torch.cat(zs, dim=0)
    if rank != 0:
        d.close()
        return
    clk = clocks.stop()
    ms, e2e_ms_max, pb_ms_max = mx
    leap_all, e2e_leap_all, div_all, pb_leap_all, passes_all = sm
    value = leap_all / (ms * 1e-3)
    zz = z.cpu().numpy().astype(np.float64)
    ess = diagnostics.effective_sample_size(zz)
    rhat = diagnostics.split_gelman_rubin(zz)
    peaks = measured_peaks()
    achieved = passes * C2_BYTES_PER_PASS / (ms * 1e-3) / 1e9                    # rank 0's kernel (every rank runs the same kernel)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_pass")
        traffic_src = "ncu --set full capture, static (profiles/traffic.json: %s), bytes per pass x passes per launch" % tj.get("source", "?")
        if traffic is not None:
            traffic = traffic * passes / max(launches // 2, 1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline("c2", X, y)
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"], "chains_total": C * world, "transitions_per_chain_per_step": T,
                   "step": "a fixed number of NUTS transitions of every chain (post warm-up); a GPU needs as many sweeps as its slowest chain needs leapfrogs",
                   "adaptation_iters_before_timing": C2_ADAPT_ITERS, "l2": "inputs_larger_than_l2 (127.8 MB swept per pass)",
                   "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
        "min_ess_per_sec": float(np.min(ess) / (ms * 1e-3)), "min_ess": float(np.min(ess)),
        "max_split_rhat": float(np.max(rhat)), "samples_per_chain": int(zz.shape[1]), "divergences": int(div_all),
        "grad_evals": int(leap_all), "mean_tree_steps": leap_all / (C * world * T * args.steps),
        "chain_slots_busy": leap_all / max(passes_all, 1.0),
        "gpu_launches": int(launches) * world,
        "e2e": {"value": e2e_leap_all / (e2e_ms_max * 1e-3), "unit": "grad-evals/s", "h2d_bytes_per_step": int(X.nbytes + y.nbytes),
                "d2h_bytes_per_step": int(d2h), "ms": e2e_ms_max, "itemised_ms_rank0": items,
                "what": f"MCMC(NUTS(LogisticRegression), num_warmup={e2e_warm}, num_samples={e2e_samples}, num_chains={C}).run from pinned "
                        f"host arrays, per GPU; all leapfrogs (warm-up + sampling) / wall"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"], "kernel": "stream_engine_kernel<7,0,false>",
                     "algorithmic_bytes_per_pass": C2_BYTES_PER_PASS, "passes_per_launch": passes / max(launches // 2, 1),
                     "us_per_pass": ms * 1e3 / max(passes, 1), "sm_cycles_per_pass": cyc_per_pass,
                     "effective_sm_clock_ghz": cyc_per_pass / (ms * 1e6 / max(passes, 1))},
        "pass_bounded": {"value": pb_leap_all / (pb_ms_max * 1e-3), "unit": "grad-evals/s", "passes_per_step_per_gpu": C2_PASSES_PER_STEP,
                         "roofline_frac_rank0": pb_passes * C2_BYTES_PER_PASS / (pb_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "what": "every GPU runs the same number of sweeps per step; chains pause mid-tree and resume (bit-identical to an unbounded run)"},
        "cpu_baseline": cpu, "clocks": clk, "per_rank": per_rank, "many_chains": many,
    }
    print(json.dumps(line))
    d.close()


# ------------------------------------------------------------------------------------------ c3 / c4 (GEMM regime)
def gemm_engine_kwargs(config):
    from numpyro_b200 import _capi
    if config == "c3":
        return dict(global_scale=_capi.SCALE_HALFCAUCHY, group_col_begin=192, group_col_end=256, tau_scale=1.0)
    return dict(likelihood=_capi.LIK_NORMAL, local_scales=1, global_scale=_capi.SCALE_HALFCAUCHY)


def run_gemm(args):
    import torch
    from numpyro_b200 import _capi, engine as eng, families, random as b2random
    from numpyro_b200.infer import MCMC, NUTS
    config, g = args.config, GEMM[args.config]
    d = Dist()
    world, rank, dev = d.world, d.rank, d.dev
    N, D, C = g["N"], g["D"], g["C"]
    X, y = make_data(config)
    Xp, yp = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    keys = b2random.split(b2random.PRNGKey(1), C * world)[rank * C:(rank + 1) * C]
    e = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=C, X=Xp, y=yp, **gemm_engine_kwargs(config))
    assert e.regime == _capi.REGIME_GEMM
    info = e.gemm_info()
    NW = 200 if config == "c3" else 500                              # the config's warm-up length (adaptation schedule)
    e.init(keys, NW)
    lf = lambda: sum(int(s.total_leapfrogs) for s in e.state()[0])
    e.run(NW, NW, fields=(), max_passes=g["warm_passes"])           # untimed: init + the first adaptation iterations
    P, W = g["passes_per_step"], max(args.warmup, 3)
    step = lambda: e.run(NW, NW, fields=(), max_passes=P)
    for _ in range(W):
        step()
    clocks = ClockSampler(d.local)
    if rank == 0:
        clocks.start()
    lf0, p0, l0 = lf(), e.pass_count, e.launch_count
    c0 = e.debug_clocks().astype(np.float64)
    d.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    d.barrier()
    ms = ev0.elapsed_time(ev1)
    leap, passes, launches = lf() - lf0, e.pass_count - p0, e.launch_count - l0
    c1 = e.debug_clocks().astype(np.float64)
    st, _ = e.state()
    iters = [int(s.i) for s in st]

    # ---- end to end through the public API: pinned host arrays -> samples on the host (a short run: the full 200/200 or
    #      500/500 schedule is minutes of GPU time, see --full)
    fam = families.HierarchicalGLM((192, 256)) if config == "c3" else families.HorseshoeRegression("normal")
    e2e_w, e2e_s = (6, 4)
    e.close()
    e2e_ms, e2e_leap, d2h, items = 0.0, 0, 0, {}
    for it in range(2):
        d.barrier()
        t0 = time.perf_counter()
        mcmc = MCMC(NUTS(fam, max_tree_depth=6), num_warmup=e2e_w, num_samples=e2e_s, num_chains=C, chain_method="vectorized", progress_bar=False)
        mcmc.run(keys, Xp, yp)
        samples = mcmc.get_samples()
        d.barrier()
        if it > 0:
            e2e_ms = (time.perf_counter() - t0) * 1e3
            e2e_leap = mcmc.total_grad_evals
            d2h = sum(v.nbytes for v in samples.values())
            items = dict(mcmc.timings)
        for s in mcmc._shards:
            s.engine.close()
        del mcmc, samples

    mx = d.reduce([ms, e2e_ms], "MAX")
    sm = d.reduce([float(leap), float(e2e_leap)], "SUM")
    per_rank = d.gather_obj({"rank": rank, "ms": ms, "passes": int(passes), "grad_evals": int(leap)})
    if rank != 0:
        d.close()
        return
    clk = clocks.stop()
    ms, e2e_ms_max = mx
    leap_all, e2e_leap_all = sm
    peaks = measured_peaks()
    tf32_peak = 0.5 * peaks["bf16_tflops_sustained"]
    achieved = 4.0 * N * D * leap / (ms * 1e-3) / 1e12               # rank 0's kernel
    cyc = (c1 - c0)
    print("tick laps chain 0 (cycles per pass): glm_finish %.0f advance %.0f" % (cyc[6] / max(passes, 1), cyc[7] / max(passes, 1)), file=sys.stderr)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(config, X, y, depth=3, rounds=3)
    line = {
        "metric": "grad_evals_per_sec", "value": leap_all / (ms * 1e-3), "unit": "grad-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 tensor cores, 3-term split, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": WORKLOADS[config], "chains_total": C * world, "passes_per_step_per_gpu": P,
                   "step": "a fixed number of gradient passes over all chains of a GPU; chains pause between passes and resume in the next step",
                   "phase": f"early warm-up (after {g['warm_passes']} untimed passes; chains at iterations {min(iters)}..{max(iters)} of {NW}); every chain is active in every pass",
                   "l2": "inputs_larger_than_l2 (tile images of X 2 x %.0f MB + betas + partial sums per pass)" % (N * max(D, 32) * 8 / 1e6),
                   "gemm": info, "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
        "grad_evals": int(leap_all), "grad_evals_per_pass": leap / max(passes, 1), "ms_per_pass": ms / max(passes, 1),
        "gpu_launches": int(launches) * world,
        "e2e": {"value": e2e_leap_all / (e2e_ms_max * 1e-3), "unit": "grad-evals/s", "h2d_bytes_per_step": int(X.nbytes + y.nbytes),
                "d2h_bytes_per_step": int(d2h), "ms": e2e_ms_max, "itemised_ms_rank0": items,
                "what": f"MCMC(NUTS(model, max_tree_depth=6), num_warmup={e2e_w}, num_samples={e2e_s}, num_chains={C}).run from pinned host arrays, "
                        f"per GPU; all leapfrogs / wall (dominated by H2D, tile-image build and the D2H of {C} chains' samples)"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                     "traffic": None, "peak_source": peaks["source"] + "; kind::tf32 peak taken as 1/2 of the measured sustained bf16 rate",
                     "kernel": "gemm_pass_kernel", "algorithmic_flops_per_grad_eval": 4.0 * N * D,
                     "split": "every product is 3 tf32 MMAs (hi*hi + lo*hi + hi*lo) for fp32 parity: the executed tensor work is 3x the "
                              "algorithmic flops, so frac <= 1/3" + ("; more than 256 columns: one forward product per row chunk, one backward product "
                                                                  "per 256-column block from the residuals kept in tensor memory" if D > 256 else ""),
                     "executed_tensor_frac": achieved * 3 / tf32_peak,
                     "frac_of_bf16_sustained": achieved / peaks["bf16_tflops_sustained"],
                     "cta0_cycles_per_pass": cyc[0] / max(passes, 1), "cta0_wait_epilogue": cyc[2] / max(cyc[0], 1),
                     "cta0_wait_tile_copies": cyc[3] / max(cyc[0], 1), "cta0_wait_own_mma": cyc[4] / max(cyc[0], 1),
                     "cta0_wait_gbeta_drain": cyc[5] / max(cyc[0], 1)},
        "cpu_baseline": cpu, "clocks": clk, "per_rank": per_rank,
    }
    print(json.dumps(line))
    d.close()


# ------------------------------------------------------------------------------------------ c5 (row-sharded)
def run_c5(args):
    """Every rank generates ITS rows on the device (102 GB never exist on the host), all ranks replicate the 64
    chains (same keys) and exchange the per-chain sums after every GEMM pass.  A step = passes_per_step passes of all chains."""
    import torch
    from numpyro_b200 import _capi, engine as eng, random as b2random
    d = Dist()
    world, rank, dev = d.world, d.rank, d.dev
    g = dict(C5)
    g["rows_per_gpu"] = int(os.environ.get("B200NUTS_C5_ROWS", g["rows_per_gpu"]))        # (smaller shards for smoke runs)
    rows, D, C = g["rows_per_gpu"], g["D"], g["C"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(5000 + rank)
    X = torch.empty((rows, D), dtype=torch.float32, device=dev)
    blk = 1 << 20
    for lo in range(0, rows, blk):                                     # (blockwise: no second 12.8 GB temporary)
        X[lo:lo + blk].normal_(generator=gen)
    beta_true = torch.from_numpy((np.random.default_rng(5).normal(size=D) * 0.1).astype(F)).to(dev)
    y = torch.empty(rows, dtype=torch.float32, device=dev)
    for lo in range(0, rows, blk):
        y[lo:lo + blk] = (torch.rand(min(blk, rows - lo), generator=gen, device=dev) < torch.sigmoid(X[lo:lo + blk] @ beta_true)).float()
    keys = b2random.split(b2random.PRNGKey(1), C)                     # replicated chains: the same keys on every rank
    kw = dict(shard_rank=rank, shard_count=world, n_rows_global=rows * world) if world > 1 else {}
    e = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, regime=_capi.REGIME_GEMM,
                   max_tree_depth_warmup=6, max_tree_depth=6, **kw)
    if world > 1:
        e.connect_shards()                                             # mailboxes opened through CUDA IPC
    info = e.gemm_info()
    NW = g["num_warmup"]
    e.init(keys, NW)
    lf = lambda: sum(int(s.total_leapfrogs) for s in e.state()[0])
    e.run(NW, NW, fields=(), max_passes=g["warm_passes"])
    P, W = g["passes_per_step"], max(args.warmup, 3)
    step = lambda: e.run(NW, NW, fields=(), max_passes=P)
    for _ in range(W):
        step()
    clocks = ClockSampler(d.local)
    if rank == 0:
        clocks.start()
    lf0, p0, l0 = lf(), e.pass_count, e.launch_count
    c0 = e.debug_clocks().astype(np.float64)
    d.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    d.barrier()
    ms = ev0.elapsed_time(ev1)
    leap, passes, launches = lf() - lf0, e.pass_count - p0, e.launch_count - l0
    c1 = e.debug_clocks().astype(np.float64)
    st, vec = e.state()
    iters = [int(s.i) for s in st]
    # every rank must hold the same chains, bit for bit (the replicas would diverge otherwise)
    sig = torch.from_numpy(np.ascontiguousarray(vec["z"])).to(dev)
    same = True
    if world > 1:
        sigs = [torch.empty_like(sig) for _ in range(world)]
        d.dist.all_gather(sigs, sig)
        same = all(torch.equal(sigs[0].view(torch.int32), s_.view(torch.int32)) for s_ in sigs)
    mx = d.reduce([ms], "MAX")
    per_rank = d.gather_obj({"rank": rank, "ms": ms, "passes": int(passes), "grad_evals": int(leap)})
    if rank != 0:
        e.close()
        d.close()
        return
    clk = clocks.stop()
    ms = mx[0]
    peaks = measured_peaks()
    bytes_per_pass = rows * D * 4 + rows * 4                            # per GPU: its rows of X and y once per pass (all 64 chains)
    achieved = passes * bytes_per_pass / (ms * 1e-3) / 1e9
    tf32_peak = 0.5 * peaks["bf16_tflops_sustained"]
    tflops = 4.0 * rows * D * leap / (ms * 1e-3) / 1e12                # per GPU, algorithmic
    cyc = c1 - c0
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 200_000
        Xc, yc = X[:n_cpu].cpu().numpy(), y[:n_cpu].cpu().numpy()
        cpu = cpu_baseline("c5", Xc, yc, depth=3, rounds=3)
        cpu["sample"] += f"; run on the first {n_cpu} rows and scaled linearly to {rows} rows (x {n_cpu / rows:.4g})"
        cpu["value_at_sample_size"] = cpu["value"]
        cpu["value"] = cpu["value"] * n_cpu / rows
    line = {
        "metric": "grad_evals_per_sec", "value": leap / (ms * 1e-3), "unit": "grad-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 tensor cores, 3-term split, fp32 accumulate)", "data": "synthetic (generated on the device)",
        "config": {"workload": WORKLOADS["c5"], "chains_total": C, "rows_per_gpu": rows, "rows_total": rows * world, "passes_per_step": P,
                   "step": "a fixed number of gradient passes of the 64 replicated chains; every GPU sweeps its own rows in every pass",
                   "scaling_note": "weak in the data: rows per GPU are fixed, the dataset grows with the number of GPUs; grad-evals/s of the "
                                   "64 chains stays flat when the exchange is hidden (value(N) / value(1), not / N, is the efficiency)",
                   "phase": f"early warm-up (after {g['warm_passes']} untimed passes; chains at iterations {min(iters)}..{max(iters)} of {NW})",
                   "l2": "inputs_larger_than_l2 (tile images of X: 2 x %.1f GB per GPU)" % (rows * D * 8 / 1e9),
                   "gemm": info, "parallelism": f"rows sharded over {world} GPU(s); all-reduce of [64 chains x (128 + 1)] fp32 per pass through peer "
                                                "mailboxes (CUDA IPC over NVLink), sums in rank order, no NCCL on the data path"},
        "grad_evals": int(leap), "grad_evals_per_pass": leap / max(passes, 1), "ms_per_pass": ms / max(passes, 1),
        "row_grad_evals_per_sec": leap * float(rows) * world / (ms * 1e-3),
        "replicas_bit_identical": bool(same),
        "gpu_launches": int(launches) * world,
        "e2e": {"value": None, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "what": "not measured: SURVEY.md 8(d) generates every rank's shard on the device (the 102 GB dataset never exists on the host)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peaks["source"], "kernel": "gemm_pass_kernel",
                     "algorithmic_bytes_per_pass_per_gpu": bytes_per_pass,
                     "image_bytes_read_per_pass_per_gpu": rows * D * 16 + rows * 4,
                     "tensor": {"achieved_tflops": tflops, "peak_tf32": tf32_peak, "frac": tflops / tf32_peak,
                                "executed_frac": tflops * 3 * 2 / tf32_peak,
                                "note": "3-term split x 128-lane chain tile half filled by 64 chains: executed tensor work = 6x algorithmic"},
                     "side": "SURVEY 8(d) ridge: the config needs 419 algorithmic TFLOP/s per GPU to stay HBM-bound; with the fp32-parity split the "
                             "pass is tensor- and image-traffic-bound (the pre-split X / X^T tile images are 4x the algorithmic bytes)",
                     "cta0_cycles_per_pass": cyc[0] / max(passes, 1), "cta0_wait_epilogue": cyc[2] / max(cyc[0], 1),
                     "cta0_wait_tile_copies": cyc[3] / max(cyc[0], 1), "cta0_wait_own_mma": cyc[4] / max(cyc[0], 1)},
        "cpu_baseline": cpu, "clocks": clk, "per_rank": per_rank,
    }
    print(json.dumps(line))
    e.close()
    d.close()


# ------------------------------------------------------------------------------------------ c1
Y8 = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0], F)
S8 = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0], F)
README_TABLE = {"mu": (4.08, 3.51), "tau": (3.96, 3.31), "theta[0]": (6.48, 5.72)}       # README.md:118-143 (1 chain, 500/1000)


def run_c1(args):
    import torch
    from numpyro_b200 import diagnostics, families, random as b2random
    from numpyro_b200.infer import MCMC, NUTS
    d = Dist()
    if d.rank != 0:
        d.close()
        return
    mk = lambda: MCMC(NUTS(families.EightSchoolsNonCentered()), num_warmup=1000, num_samples=1000, num_chains=4,
                      chain_method="vectorized", progress_bar=False)
    W = max(args.warmup, 3)
    for k in range(W):
        m = mk()
        m.run(b2random.PRNGKey(100 + k), 8, S8, y=Y8)
    clocks = ClockSampler(d.local)
    clocks.start()
    leap, t_ms, last = 0, 0.0, None
    for k in range(args.steps):
        m = mk()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.run(b2random.PRNGKey(k), 8, S8, y=Y8, extra_fields=("num_steps",))
        s = m.get_samples(group_by_chain=True)
        t_ms += (time.perf_counter() - t0) * 1e3
        leap += m.total_grad_evals
        last = (m, s)
    clk = clocks.stop()
    m, s = last
    summ = diagnostics.summary({k: v for k, v in s.items() if k in ("mu", "tau", "theta")}, group_by_chain=True)
    table = {"mu": [float(summ["mu"][k]) for k in ("mean", "std", "n_eff", "r_hat")],
             "tau": [float(summ["tau"][k]) for k in ("mean", "std", "n_eff", "r_hat")],
             "theta[0]": [float(np.ravel(summ["theta"][k])[0]) for k in ("mean", "std", "n_eff", "r_hat")]}
    ok = all(abs(table[k][0] - README_TABLE[k][0]) < 4 * README_TABLE[k][1] / np.sqrt(max(table[k][2], 1.0)) + 0.5 and table[k][3] < 1.01 + 0.01
             for k in README_TABLE)
    transitions = 4 * 2000 * args.steps
    cpu = None if args.no_cpu_baseline else c1_cpu()
    line = {"metric": "grad_evals_per_sec", "value": leap / (t_ms * 1e-3), "unit": "grad-evals/s", "n_gpus": 1, "steps": args.steps,
            "warmup": W, "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "README.md:35-37 (J=8, y, sigma)",
            "config": {"workload": WORKLOADS["c1"], "step": "one whole MCMC.run (init + 1000 warm-up + 1000 samples, 4 chains) through the public API, host arrays in, samples out",
                       "l2": "not applicable (latency-bound: D = 10, one warp per chain, the whole run is one launch)"},
            "transitions_per_sec": transitions / (t_ms * 1e-3), "grad_evals": leap,
            "posterior": {"columns": ["mean", "std", "n_eff", "r_hat"], "table": table, "readme": README_TABLE,
                          "divergences": int(m.get_extra_fields()["diverging"].sum()), "matches_readme": bool(ok)},
            "gpu_launches": 6 * args.steps,
            "e2e": {"value": leap / (t_ms * 1e-3), "unit": "grad-evals/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": int(sum(v.nbytes for v in s.values())),
                    "what": "identical to `value`: the step already is the public-API call with host buffers"},
            "roofline": {"bound": "hbm", "achieved": None, "peak": measured_peaks()["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "latency-bound by construction (SURVEY.md 8(d): report transitions/s only)"},
            "cpu_baseline": cpu, "clocks": clk}
    print(json.dumps(line))
    d.close()


def c1_cpu(transitions=150):
    """The oracle port on one host core: one chain, `transitions` NUTS transitions from a fixed start."""
    from oracle import chain, families, prng
    fam = families.EightSchools(S8, Y8)
    kern = chain.Kernel(fam.potential_and_grad)
    t0 = time.perf_counter()
    res, _ = chain.run_chain(kern, fam, prng.key(0), transitions // 2, transitions - transitions // 2, fields=("num_steps",))
    # (warm-up trees are not collected: count them through a second, identical run's trace is not worth it -- rate by samples)
    dt = time.perf_counter() - t0
    leaps = float(np.sum(res["num_steps"])) * 2.0                     # warm-up half assumed like the sampling half
    return {"value": leaps / dt, "unit": "grad-evals/s", "cores": 1, "kind": "port",
            "sample": f"1 process, 1 chain, {transitions} oracle NUTS transitions (pure NumPy, fp64 potential) in {dt:.1f} s; "
                      "leapfrogs of the collected half doubled"}


def run_c1_reference(args):
    cpu = c1_cpu(300)
    line = {"impl": "reference", "metric": "grad_evals_per_sec", "value": cpu["value"], "unit": "grad-evals/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "README.md:35-37", "config": {"workload": WORKLOADS["c1"], "arm": "restated reference (oracle port)"},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-many-chains", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        {"c1": run_c1, "c2": run_c2, "c3": run_gemm, "c4": run_gemm, "c5": run_c5}[args.config](args)


if __name__ == "__main__":
    main()
