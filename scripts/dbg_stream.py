import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numpyro_b200 import _capi, engine as eng
from oracle import prng
F = np.float32
rng = np.random.default_rng(4)
N, D, C = 6000, 7, 3
X = rng.normal(size=(N, D)).astype(F)
y = (rng.uniform(size=N) < 0.5).astype(F)
e = eng.Engine(family=_capi.FAMILY_GLM, num_chains=C, X=X, y=y, regime=_capi.REGIME_STREAM, max_tree_depth_warmup=3, max_tree_depth=3)
e.init(prng.split(prng.key(7), C), 4)
out = e.run(6, 4)
torch.cuda.synchronize()
print(out["num_steps"], out["z"][0, :, :3])
