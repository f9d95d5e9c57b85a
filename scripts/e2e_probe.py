"""Where does the end-to-end time of bench.py's e2e leg go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from numpyro_b200 import families, random as b2random
from numpyro_b200.infer import MCMC, NUTS
X, y = bench.make_data()
Xp, yp = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
keys = b2random.split(b2random.PRNGKey(1), 8)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mcmc = MCMC(NUTS(families.LogisticRegression()), num_warmup=30, num_samples=10, num_chains=8, chain_method="vectorized", progress_bar=False)
    mcmc._ensure_engines((Xp, yp), {})
    torch.cuda.synchronize(); t1 = time.perf_counter()
    e = mcmc._shards[0].engine
    p0 = e.pass_count
    mcmc._shards_keep = True
    # the rest of run() re-creates the engines when fresh=True, so time the pieces by hand
    e.init(keys, 30); torch.cuda.synchronize(); t2 = time.perf_counter()
    out = e.run(40, 30, fields=("z", "diverging", "num_steps")); torch.cuda.synchronize(); t3 = time.perf_counter()
    con = e.constrain(out["z"]); host = {k: v.cpu().numpy() for k, v in out.items()}; st, vec = e.state(); t4 = time.perf_counter()
    leap = sum(int(s.total_leapfrogs) for s in st)
    print(f"iter {it}: create {1e3*(t1-t0):.1f} ms, init {1e3*(t2-t1):.1f} ms, run {1e3*(t3-t2):.1f} ms ({e.pass_count - p0} passes, {leap} leapfrogs, {leap/max(e.pass_count-p0,1):.2f}/pass), fetch {1e3*(t4-t3):.1f} ms; total {1e3*(t4-t0):.1f} ms -> {leap/(t4-t0):,.0f} grad-evals/s")
    e.close()
