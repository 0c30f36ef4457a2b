"""N > 1 host logic on CPU: world_size-2 gloo jobs exercising the partition arithmetic, the
NCCL-id exchange, the reassembly of sharded results and the fixed-rank-order summation,
with the oracle standing in for the per-rank device work."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from sqlp_b200 import dist as D
from tests.helpers import load_instance, sample_instance_values


def test_partition_arithmetic_matches_brute_force():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 127, 128, 129, 1000, 128 * world, 128 * world + 5, 5000):
            g = np.arange(n)
            own = D.owner_of(g, world)
            seen = 0
            for r in range(world):
                mine = g[own == r]
                assert D.local_count(n, r, world) == len(mine)
                assert np.array_equal(D.local_of(mine, world), np.arange(len(mine)))
                assert np.array_equal(D.global_of(r, np.arange(len(mine)), world), mine)
                assert np.array_equal(D.owned_ordinals(n, r, world), mine)
                seen += len(mine)
            assert seen == n
            if n:   # appended scenarios keep shards within one tile of each other
                counts = [D.local_count(n, r, world) for r in range(world)]
                assert max(counts) - min(counts) <= D.TILE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        uid = D.exchange_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))

        P, z = load_instance("baa99-20")
        N = 700
        vals = sample_instance_values(z, N)
        w = 0.5 + O.u01(4, np.arange(N))
        x = z["x_alt"]
        pool = z["pool"]
        mine = D.owned_ordinals(N, rank, world)
        W = 0.0
        for v in w:
            W += float(v)
        # per-rank work (what the device does on its shard): argmax + partial weighted sums
        mv, mi = O.argmax_procedure(P, vals[mine], x, pool)
        part = O.build_sasa_cut(P, vals[mine], w[mine], x, pool, total_weight=W)
        local = np.concatenate([[part["alpha"]], part["beta"], [part["val"]]])
        total = D.ordered_rank_sum(local)
        gmv = D.gather_scenario_results(mv, N)
        gmi = D.gather_scenario_results(mi, N)
        full = O.build_sasa_cut(P, vals, w, x, pool)
        ok = (np.array_equal(gmi, full["max_idx"]) and np.array_equal(gmv, full["max_val"])
              and abs(total[0] - full["alpha"]) <= 1e-12 * abs(full["alpha"])
              and np.allclose(total[1:-1], full["beta"], rtol=1e-12, atol=1e-9)
              and abs(total[-1] - full["val"]) <= 1e-12 * abs(full["val"]))
        q.put((rank, bool(ok), total.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharding_and_ordered_sum():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]          # every rank holds the same bits
