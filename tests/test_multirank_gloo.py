"""World-size-2 CPU test (gloo) of the chain-sharded multi-GPU host logic: rank r owns chains
[r*C/W, (r+1)*C/W), keys are rows of split(key, C) (mcmc.py:670-671) so results are independent of W,
and the only communication is the final all_gather of the collected arrays."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import prng


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, C, S, D, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from numpyro_b200 import families
        from numpyro_b200.infer import MCMC, NUTS
        mcmc = MCMC(NUTS(families.LogisticRegression()), num_warmup=10, num_samples=S, num_chains=C,
                    chain_method="parallel", progress_bar=False)
        assert mcmc._dist
        (shard,) = mcmc._plan_shards()
        per = C // world
        assert (shard.lo, shard.hi) == (rank * per, (rank + 1) * per)
        keys = prng.split(prng.key(1), C)[shard.lo:shard.hi]
        # stand-in for the per-rank engine output: a deterministic function of the chain's key
        z = np.stack([np.outer(np.arange(S), np.ones(D)) + float(k[0] % 1000) for k in keys]).astype(np.float32)
        host = {"z": z, "diverging": (z[..., 0] % 2).astype(np.int32)}
        full = mcmc._all_gather(host)
        q.put((rank, full["z"].shape, float(full["z"].sum()), full["z"][:, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_chain_sharding_and_final_gather_world2():
    C, S, D, world = 4, 5, 3, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, C, S, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    keys = prng.split(prng.key(1), C)
    want_first = [float(k[0] % 1000) for k in keys]
    for rank, shape, total, first in res:
        assert shape == (C, S, D)
        assert first == want_first                      # chain order == key order, identical on every rank
    assert res[0][2] == res[1][2]


def test_ragged_chain_count_is_rejected():
    from numpyro_b200 import families
    from numpyro_b200.infer import MCMC, NUTS
    m = MCMC(NUTS(families.LogisticRegression()), num_warmup=1, num_samples=1, num_chains=3, chain_method="vectorized")
    (s,) = m._plan_shards()
    assert (s.lo, s.hi) == (0, 3)
