"""Tick lap table of the covtype-shaped streaming run in its post-warm-up phase (dev build with -DB2_TICK_LAPS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numpyro_b200 import _capi, engine as eng
from oracle import prng
F = np.float32
N, D, C = 581012, 54, 8
rng = np.random.default_rng(1)
X = rng.standard_normal(size=(N, D), dtype=F)
beta = (rng.normal(size=D) * 0.3).astype(F)
y = (rng.uniform(size=N) < 1 / (1 + np.exp(-(X @ beta)))).astype(F)
names = ["wait_beta", "sweep", "cta_reduce+publish", "wait_partials", "xcta_reduce", "total", "tick_busy", "beta_frags",
         "tick_finish", "tick_advance", "tick_publish"]
W = int(sys.argv[1]) if len(sys.argv) > 1 else 300
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y)
e.init(prng.split(prng.key(1), C), W)
os.environ.pop("B200NUTS_DEBUG_TICK", None)
e.run(W, W, fields=())
torch.cuda.synchronize()
os.environ["B200NUTS_DEBUG_TICK"] = "1"; os.environ["B200NUTS_DEBUG_CTA"] = "1"
p0 = e.pass_count
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); out = e.run(W + 200, W, fields=("num_steps",)); t1.record(); torch.cuda.synchronize()
passes = e.pass_count - p0
dbg = e.debug_clocks().astype(np.float64)
print("post warm-up: passes", passes, "grad-evals", int(out["num_steps"].sum()), "us/pass %.2f" % (t0.elapsed_time(t1) * 1e3 / passes),
      {n: round(dbg[i] / passes) for i, n in enumerate(names)})
