"""A short multi-pass launch of the streaming engine on the covtype shape for `ncu --set full`:
launch 0 = adaptation (60 transitions), launch 1 = the profiled one (12 transitions, ~100-150 passes)."""
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
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y)
e.init(prng.split(prng.key(1), C), 60)
e.run(60, 60, fields=())
p0 = e.pass_count
out = e.run(72, 60, fields=("num_steps",))
torch.cuda.synchronize()
print("profiled launch: passes", e.pass_count - p0, "grad evals", int(out["num_steps"].sum().item()))
