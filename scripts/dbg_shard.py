import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from scipy.special import gammaln
os.environ["B200NUTS_GRID"] = "74"
import test_gpu_rowshard as t
from numpyro_b200 import _capi
F = np.float32
N, D, C = 30011, 54, 5
X, y = t._data(N, D, 3, "poisson")
rk = t.Ranks(X, y, C, [0, 17003, N], (0, 0), likelihood=_capi.LIK_POISSON_LOG)
rng = np.random.default_rng(0)
z = (rng.normal(size=(C, D)) * 0.3).astype(F)
res = rk.each(lambda e: tuple(x.cpu().numpy() for x in e.potential_and_grad(z)))
print("U rank0", res[0][0]); print("U rank1", res[1][0])
eta = X.astype(np.float64) @ z.astype(np.float64).T
for lo, hi in ((0, 17003), (17003, N)):
    nll = (np.exp(eta[lo:hi]) - y[lo:hi, None] * eta[lo:hi]).sum(0)
    print("rows", lo, hi, "nll", nll, "const", gammaln(y[lo:hi].astype(np.float64) + 1).sum())
prior = 0.5 * (z.astype(np.float64) ** 2).sum(1) + D * 0.9189385332046727
print("prior", prior)
res = rk.each(lambda e: tuple(x.cpu().numpy() for x in e.potential_and_grad(z)))
print("again U rank0", res[0][0]); print("again U rank1", res[1][0])
