#!/usr/bin/env python
"""Headline benchmark: gradient evaluations per second (and min-ESS per second) of NUTS on the
covtype-shaped Bayesian logistic regression (BASELINE.json configs[1]: N = 581012, D = 54 fp32,
8 chains per GPU; synthetic data, same shapes as examples/covtype.py).

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA engine)
    python bench.py --impl reference --gpus N ...          # CPU arm: the oracle port on host cores

A *step* is one pass-bounded call of the engine: PASSES_PER_STEP sweeps of X per GPU, post warm-up (adapted step size /
mass matrix).  Every chain advances its NUTS transitions as far as those sweeps carry it and pauses wherever it is in its
tree; the next step resumes it (bit-identical to an unbounded run, tests/test_gpu_parity.py).  Every GPU therefore does the
same amount of work per step, whatever tree depths its chains happen to have adapted to.  Adaptation runs before the timed
region as setup.
`value` is measured with the dataset resident in HBM; `e2e` runs the public ``MCMC`` API from pinned
host buffers (H2D of X and y, init, warm-up, sampling, D2H of the samples) inside the timed region.
Multi-GPU (torchrun, one rank per GPU): chains shard across ranks with no data-path collective
(weak scaling, 8 chains per GPU, dataset replicated); time = max over ranks.
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

N_ROWS, N_COLS, CHAINS_PER_GPU = 581012, 54, 8
BYTES_PER_PASS = N_ROWS * N_COLS * 4 + N_ROWS * 4          # one sweep of X and y serves every chain
PASSES_PER_STEP = 3000          # sweeps of X per step and GPU (about what 400 transitions of 8 chains need)
ADAPT_ITERS = 600
WORKLOAD = ("configs[1]: covtype-shaped Bayesian logistic regression NUTS "
            "(N=581012, D=54 fp32, synthetic, 8 chains per GPU, max_tree_depth=10)")


def make_data(seed=1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal(size=(N_ROWS, N_COLS), dtype=np.float32)
    X = (X - X.mean(0)) / X.std(0)                        # column-standardised as examples/covtype.py:48
    beta = (rng.normal(size=N_COLS) * 0.3).astype(np.float32)
    y = (rng.uniform(size=N_ROWS) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float32)
    return np.ascontiguousarray(X, np.float32), y


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


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


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_DATA = {}          # dataset inherited by the forked workers (never pickled)


def _cpu_worker(args):
    """One oracle NUTS transition (tree depth capped so it finishes) on one host core."""
    z0, step, imm, key, depth = args
    X, y = _CPU_DATA["X"], _CPU_DATA["y"]
    from oracle import chain, families, prng
    from oracle.tree import build_tree
    fam = families.logistic_regression(X, y)
    pot = fam.potential_and_grad_f32
    F = np.float32
    u, g = pot(z0)                                           # warm caches / build the fp32 copies
    r = (1.0 / np.sqrt(imm) * prng.normal(key, z0.shape[0])).astype(F)
    t0 = time.perf_counter()
    tree = build_tree(pot, imm, F(step), key, z0, r, u, g, depth)
    return tree.num_proposals, time.perf_counter() - t0


def cpu_sample(X, y, z, step, imm, cores, depth=5, rounds=1):
    """Bounded CPU sample: `cores` processes x `rounds` oracle transitions (depth cap `depth`)."""
    import multiprocessing as mp
    from oracle import prng
    keys = prng.split(prng.key(123), cores * rounds)
    _CPU_DATA["X"], _CPU_DATA["y"] = X, y
    jobs = [(z[i % z.shape[0]], float(step[i % len(step)]), imm[i % imm.shape[0]], keys[i], depth)
            for i in range(cores * rounds)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    leap = sum(r[0] for r in res)
    busy = max(r[1] for r in res) if rounds == 1 else wall
    return leap, busy, wall


def default_start(D):
    """Start point for the CPU arm when no adapted GPU state is available: near the mode scale."""
    rng = np.random.default_rng(7)
    z = (rng.normal(size=(CHAINS_PER_GPU, D)) * 0.05).astype(np.float32)
    return z, np.full(CHAINS_PER_GPU, 0.02, np.float32), np.full((CHAINS_PER_GPU, D), 1e-4, np.float32)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y = make_data()
    cores = os.cpu_count() or 1
    z, step, imm = default_start(N_COLS)
    for _ in range(args.warmup):
        cpu_sample(X, y, z, step, imm, cores, depth=2)
    leap_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        leap, busy, wall = cpu_sample(X, y, z, step, imm, cores, depth=5)
        leap_tot += leap
        t_tot += busy
    value = leap_tot / t_tot
    sample = (f"{cores} processes x 1 oracle NUTS transition each per step (tree depth capped at 5 = <=31 "
              f"leapfrogs, fp32 BLAS potential), {args.steps} steps; process start-up excluded")
    line = {"impl": "reference", "metric": "grad_evals_per_sec", "value": value, "unit": "grad-evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arm": "restated reference (oracle port; jax/numpyro are not installable here)"},
            "cpu_baseline": {"value": value, "unit": "grad-evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from numpyro_b200 import _capi, diagnostics, engine as eng, families, random as b2random
    from numpyro_b200.infer import MCMC, NUTS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, y = make_data()
    Xp, yp = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    C_total = CHAINS_PER_GPU * world
    keys = b2random.split(b2random.PRNGKey(1), C_total)[rank * CHAINS_PER_GPU:(rank + 1) * CHAINS_PER_GPU]

    # ---- setup (untimed): dataset to HBM, chain init, warm-up adaptation
    e = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=CHAINS_PER_GPU, X=Xp, y=yp)
    assert e.regime == _capi.REGIME_STREAM
    e.init(keys, ADAPT_ITERS)
    e.run(ADAPT_ITERS, ADAPT_ITERS, fields=())
    fields = ("z", "num_steps", "diverging")
    n_steps_total = max(args.warmup, 3) + args.steps + 2
    lower, upper = ADAPT_ITERS, ADAPT_ITERS + n_steps_total * (PASSES_PER_STEP // 2)     # window no chain can outrun
    out = None

    def step():
        nonlocal out
        out = e.run(upper, lower, fields=fields, max_passes=PASSES_PER_STEP, out=out)
        return out

    def progress():
        st_, _ = e.state()
        return np.array([s_.i for s_ in st_]), np.array([int(s_.total_leapfrogs) for s_ in st_])

    for _ in range(max(args.warmup, 3)):
        step()
    # Optional extra untimed load (off by default).  Back-to-back runs of identical work on one box of this pool scatter
    # between 38.6 and 50 us per pass (clock / power state of the node); 20 s of extra load before timing did not remove the
    # scatter (40.2 then 50.0 us), so the default stays at the contract's W warm-up steps.
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < args.gpu_warm_seconds:
        step()
        torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    it0, lf0 = progress()
    barrier()
    l0, p0 = e.launch_count, e.pass_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches, passes = e.launch_count - l0, e.pass_count - p0
    it1, lf1 = progress()
    # SM cycles per pass of the last timed launch (clock64 on CTA 0): with us_per_pass it gives the SM clock the run really had
    cyc_per_pass = float(e.debug_clocks()[5]) / max(passes / max(args.steps, 1), 1.0)
    leap = int((lf1 - lf0).sum())                                     # leapfrogs of the timed region, partial trees included
    # samples: the transitions every chain (of every rank) completed inside the timed region
    rng_t = torch.tensor([float(it0.max()), float(-it1.min())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rng_t, op=dist.ReduceOp.MAX)
    t_lo, t_hi = int(rng_t[0].item()), int(-rng_t[1].item())
    z = out["z"][:, t_lo - lower:t_hi - lower].contiguous()           # [C, T, D]
    diverging = int(out["diverging"][:, t_lo - lower:t_hi - lower].sum().item())
    transitions_done = int((it1 - it0).sum())
    st, vec = e.state()

    # ---- end to end through the public API (pinned host inputs -> samples on the host)
    e2e_steps = max(1, min(args.steps, 2))
    e2e_warm, e2e_samples = 30, 10
    model = families.LogisticRegression()
    e2e_leap, e2e_ms, d2h = 0, 0.0, 0
    for it in range(1 + e2e_steps):                                   # first iteration is a warm-up
        barrier()
        t0 = time.perf_counter()
        mcmc = MCMC(NUTS(model), num_warmup=e2e_warm, num_samples=e2e_samples, num_chains=CHAINS_PER_GPU,
                    chain_method="vectorized", progress_bar=False)
        mcmc.run(keys, Xp, yp, extra_fields=("num_steps",))
        samples = mcmc.get_samples()
        barrier()
        dt = (time.perf_counter() - t0) * 1e3
        if it > 0:
            e2e_ms += dt
            e2e_leap += mcmc.total_grad_evals
            d2h = sum(v.nbytes for v in samples.values()) + mcmc.get_extra_fields()["num_steps"].nbytes
        for s in mcmc._shards:
            s.engine.close()

    # ---- secondary measurement (not the headline config): the same dataset with 32 chains per GPU.  Passes then rotate over
    #      4 chain groups and the owners' ticks hide behind the other groups' sweeps.
    many = None
    if not args.no_many_chains:
        CM = 32
        keys_m = b2random.split(b2random.PRNGKey(2), CM * world)[rank * CM:(rank + 1) * CM]
        em = eng.Engine(device=dev, family=_capi.FAMILY_GLM, num_chains=CM, X=e.X, y=e.y)
        em.init(keys_m, 300)
        em.run(300, 300, fields=())
        em.run(350, 300, fields=("num_steps",))
        barrier()
        pm0 = em.pass_count
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        om = em.run(550, 350, fields=("num_steps",))
        m1.record()
        barrier()
        mms, mpasses, mleap = m0.elapsed_time(m1), em.pass_count - pm0, int(om["num_steps"].sum().item())
        many = {"chains_per_gpu": CM, "transitions": 200, "grad_evals_per_sec_this_gpu": mleap / (mms * 1e-3),
                "us_per_pass": mms * 1e3 / max(mpasses, 1), "grad_evals_per_pass": mleap / max(mpasses, 1),
                "roofline_frac": mpasses * BYTES_PER_PASS / (mms * 1e-3) / 1e9 / measured_peaks()[0]}
        em.close()

    # ---- reduce over ranks
    stats = torch.tensor([ms, float(leap), e2e_ms, float(e2e_leap), float(diverging), float(transitions_done)], dtype=torch.float64, device=dev)
    per_rank = [{"rank": rank, "ms": ms, "passes": int(passes), "grad_evals": leap,
                 "step_size": [round(float(s_.step_size), 5) for s_ in st]}]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = mx[0].item(), mx[2].item()
        leap, e2e_leap, diverging = int(sm[1].item()), int(sm[3].item()), int(sm[4].item())
        transitions_done = int(sm[5].item())
        zs = [torch.empty_like(z) for _ in range(world)]
        dist.all_gather(zs, z)                                        # final gather for the diagnostics only
        z = torch.cat(zs, dim=0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clk = clocks.stop()
    transitions_done_total = transitions_done
    value = leap / (ms * 1e-3)
    zz = z.cpu().numpy().astype(np.float64)
    ess = diagnostics.effective_sample_size(zz)
    rhat = diagnostics.split_gelman_rubin(zz)
    peak, peak_src = measured_peaks()
    achieved = passes * BYTES_PER_PASS / (ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_pass")
        if traffic is not None:
            traffic = traffic * passes / max(launches // 2, 1)
    # CPU baseline beside it (rank 0, N = 1 only): the oracle port from the adapted GPU state
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        zc, stepc, immc = vec["z"], np.array([s.step_size for s in st], np.float32), vec["inverse_mass_matrix"]
        cpu_sample(X, y, zc, stepc, immc, cores, depth=1)             # warm-up (fork + BLAS)
        cl, cb, cwall = cpu_sample(X, y, zc, stepc, immc, cores, depth=5)
        cpu = {"value": cl / cb, "unit": "grad-evals/s", "cores": cores, "kind": "port",
               "sample": f"{cores} processes x 1 oracle NUTS transition from the adapted state, tree depth capped at 5 "
                         f"(<=31 leapfrogs each, {cl} in total), fp32 BLAS potential; {cwall:.1f} s wall"}
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_total": C_total, "passes_per_step_per_gpu": PASSES_PER_STEP,
                   "step": "PASSES_PER_STEP sweeps of X per GPU; chains pause mid-tree and resume in the next step",
                   "untimed_gpu_warm_seconds": args.gpu_warm_seconds,
                   "adaptation_iters_before_timing": ADAPT_ITERS, "l2": "inputs_larger_than_l2 (127.8 MB swept per pass)",
                   "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
        "min_ess_per_sec": float(np.min(ess) / (ms * 1e-3)), "min_ess": float(np.min(ess)),
        "max_split_rhat": float(np.max(rhat)), "samples_per_chain": int(zz.shape[1]), "divergences": diverging,
        "grad_evals": leap, "transitions_completed": transitions_done_total, "mean_tree_steps": leap / max(transitions_done_total, 1),
        "gpu_launches": int(launches) * world,
        "e2e": {"value": e2e_leap / (e2e_ms * 1e-3), "unit": "grad-evals/s", "h2d_bytes_per_step": int(X.nbytes + y.nbytes),
                "d2h_bytes_per_step": int(d2h),
                "what": f"MCMC(NUTS(LogisticRegression), num_warmup={e2e_warm}, num_samples={e2e_samples}).run from pinned "
                        f"host arrays, per GPU; all leapfrogs (warm-up + sampling) / wall"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "stream_engine_kernel<7>",
                     "algorithmic_bytes_per_pass": BYTES_PER_PASS, "passes_per_launch": passes / max(launches // 2, 1),
                     "us_per_pass": ms * 1e3 / max(passes, 1), "sm_cycles_per_pass": cyc_per_pass,
                     "effective_sm_clock_ghz": cyc_per_pass / (ms * 1e6 / max(passes, 1))},
        "cpu_baseline": cpu, "clocks": clk, "per_rank": per_rank, "many_chains": many,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-many-chains", action="store_true")
    ap.add_argument("--gpu-warm-seconds", type=float, default=0.0, help="extra untimed load before the timed steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
