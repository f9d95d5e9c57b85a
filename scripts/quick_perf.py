"""Quick throughput probe of the streaming engine at the covtype shape (not the bench)."""
import sys, os, time
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
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y)
print("regime", e.regime)
z = (rng.normal(size=(C, D)) * 0.1).astype(F)
for _ in range(3):
    e.potential_and_grad(z)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(20):
    e.potential_and_grad(z)
t1.record(); torch.cuda.synchronize()
print("single-pass launch (incl. launch overhead): %.1f us" % (t0.elapsed_time(t1) / 20 * 1e3))
e.init(prng.split(prng.key(1), C), 60)
t0.record()
out = e.run(60, 60)
t1.record(); torch.cuda.synchronize()
st, _ = e.state()
lf = sum(int(s.total_leapfrogs) for s in st)
mx = max(int(s.total_leapfrogs) for s in st)
ms = t0.elapsed_time(t1)
print("warmup 60 iters: %.1f ms, leapfrogs total %d max-chain %d -> %.1f us/pass, %.0f grad-evals/s" % (ms, lf, mx, ms * 1e3 / mx, lf / ms * 1e3))
print("step sizes", [round(s.step_size, 5) for s in st])
t0.record()
out = e.run(100, 60, fields=("z", "num_steps", "diverging"))
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
ns = out["num_steps"].cpu().numpy()
print("sample 40 iters: %.1f ms, leapfrogs %d (per chain %s) -> %.0f grad-evals/s; us/pass(max chain) %.2f; HBM-roofline frac %.3f" % (
    ms, ns.sum(), ns.sum(1), ns.sum() / ms * 1e3, ms * 1e3 / ns.sum(1).max(), (ns.sum(1).max() * 127822640 / (ms * 1e-3)) / 6545.6e9))
print("divergences", int(out["diverging"].sum()))
dbg = e.debug_clocks().astype(np.float64)
names = ["wait_beta", "sweep", "cta_reduce+publish", "wait_partials", "xcta_reduce", "total", "tick_busy", "beta_frags", "tick_finish", "tick_advance", "tick_publish"]
print("CTA0 clock breakdown per pass:", {n: round(dbg[i] / max(e.pass_count and ns.sum(1).max(), 1), 0) for i, n in enumerate(names)})
