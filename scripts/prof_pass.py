"""One covtype-shaped sweep (potential_and_grad through the streaming engine) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numpyro_b200 import _capi, engine as eng
F = np.float32
N, D, C = 581012, 54, 8
rng = np.random.default_rng(1)
X = rng.standard_normal(size=(N, D), dtype=F)
y = (rng.uniform(size=N) < 0.5).astype(F)
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y)
z = (rng.normal(size=(C, D)) * 0.1).astype(F)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    e.potential_and_grad(z)
torch.cuda.synchronize()
