"""Throughput of the streaming engine on the covtype shape as a function of the number of chains (groups of 8 rotate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numpyro_b200 import _capi, engine as eng
from oracle import prng
F = np.float32
N, D = 581012, 54
rng = np.random.default_rng(1)
X = rng.standard_normal(size=(N, D), dtype=F)
beta = (rng.normal(size=D) * 0.3).astype(F)
y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
for C in [int(a) for a in sys.argv[1:]] or [8, 16, 32]:
    e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=Xd, y=yd, max_tree_depth=6, max_tree_depth_warmup=6)
    e.init(prng.split(prng.key(1), C), 60)
    e.run(60, 60, fields=())
    torch.cuda.synchronize()
    p0 = e.pass_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); out = e.run(120, 60, fields=("num_steps",)); t1.record(); torch.cuda.synchronize()
    passes = e.pass_count - p0; ms = t0.elapsed_time(t1); leap = int(out["num_steps"].sum().item())
    print(f"chains {C:3d}: passes {passes} us/pass {ms * 1e3 / passes:.2f} grad-evals/s {leap / ms * 1e3:,.0f} "
          f"(evals/pass {leap / passes:.2f}) achieved {passes * 127822640 / ms / 1e6:.0f} GB/s = {passes * 127822640 / ms / 1e6 / 6545.6:.1%} of HBM roofline", flush=True)
    names = ["wait_beta", "sweep", "cta_reduce+publish", "wait_partials", "xcta_reduce", "total", "tick_busy", "beta_frags"]
    dbg = e.debug_clocks().astype(np.float64)
    print("      cta0 cycles/pass:", {n: round(dbg[i] / passes) for i, n in enumerate(names)}, flush=True)
    e.close()
