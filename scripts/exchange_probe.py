"""Exchange floor: per-pass time of the streaming engine when the sweep is negligible (few rows), covtype columns/chains."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numpyro_b200 import _capi, engine as eng
from oracle import prng
F = np.float32
N, D, C = 148 * 15 * 16, (int(sys.argv[1]) if len(sys.argv) > 1 else 54), 8
rng = np.random.default_rng(1)
X = rng.standard_normal(size=(N, D), dtype=F)
beta = (rng.normal(size=D) * 0.3).astype(F)
y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
names = ["wait_beta", "sweep", "cta_reduce+publish", "wait_partials", "xcta_reduce", "total", "tick_busy", "beta_frags",
         "tick_finish", "tick_advance", "tick_publish"]
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, max_tree_depth=6, max_tree_depth_warmup=6, regime=_capi.REGIME_STREAM)
e.init(prng.split(prng.key(1), C), 20)
e.run(20, 20, fields=())
torch.cuda.synchronize()
p0 = e.pass_count
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); e.run(100, 20, fields=("num_steps",)); t1.record(); torch.cuda.synchronize()
passes = e.pass_count - p0
dbg = e.debug_clocks().astype(np.float64)
print("N", N, "passes", passes, "us/pass %.2f" % (t0.elapsed_time(t1) * 1e3 / passes), {n: round(dbg[i] / passes) for i, n in enumerate(names)})
