"""torchrun worker of tests/test_gpu_rowshard.py::test_sharded_processes_over_cuda_ipc: one rank per GPU, mailboxes
opened through CUDA IPC, chains replicated; rank 0 checks that every rank produced the same bits."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpyro_b200 import _capi, engine as eng          # noqa: E402
from oracle import families, prng                       # noqa: E402

F = np.float32
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ.setdefault("B200NUTS_WATCHDOG_S", "120")
GEMM = len(sys.argv) > 1 and sys.argv[1] == "gemm"          # config-5 shape of the chain state: 64 chains x 128 columns
N, D, C = (100003, 128, 64) if GEMM else (200003, 54, 8)
rng = np.random.default_rng(5)
X = rng.normal(size=(N, D)).astype(F)
y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ (rng.normal(size=D) * 0.3)))).astype(F)
cuts = [N * r // world for r in range(world + 1)]
e = eng.Engine(device=f"cuda:{local}", family=_capi.FAMILY_GLM, num_chains=C, X=X[cuts[rank]:cuts[rank + 1]],
               y=y[cuts[rank]:cuts[rank + 1]], regime=_capi.REGIME_GEMM if GEMM else _capi.REGIME_STREAM, shard_rank=rank, shard_count=world,
               n_rows_global=N, max_tree_depth=6, max_tree_depth_warmup=6)
e.connect_shards()
z = (rng.normal(size=(C, D)) * 0.2).astype(F)
U, g = e.potential_and_grad(z)
W0, S0 = (10, 6) if GEMM else (30, 20)
e.init(prng.split(prng.key(3), C), W0)
out = e.run(W0 + S0, W0, fields=("z", "num_steps"))
torch.cuda.synchronize()
mine = torch.cat([U.flatten(), g.flatten(), out["z"].flatten(), out["num_steps"].flatten().float()])
allr = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allr, mine)
if rank == 0:
    for r in range(1, world):
        assert torch.equal(allr[0].view(torch.int32), allr[r].view(torch.int32)), f"rank {r} differs from rank 0"
    fam = families.logistic_regression(X, y)
    u64, g64 = fam.potential64(z[2].astype(np.float64))
    assert abs(U[2].item() - u64) <= 1e-5 * abs(u64)
    assert np.allclose(g[2].cpu().numpy(), g64, rtol=1e-5, atol=1e-5 * np.abs(g64).max())
    print("ROWSHARD_OK passes", e.pass_count, "grad evals", int(out["num_steps"].sum().item()))
e.close()
dist.destroy_process_group()
