"""Scenario sharding across the GPUs of one box (one process per GPU).

The path shards naturally: given (x, pool) every scenario's argmax is
independent (reference ``subprob.jl:148-166``) and the cut is a weighted sum over scenarios
(``epigraph.jl:134-143``).  Scenario ``g`` of an epigraph lives on rank ``(g // 128) % world``
(block-cyclic in 128-scenario tiles so appended scenarios keep the shards balanced); the
dual-vertex pool is replicated.  The data path has two collectives, both inside
``libsqlp_b200.so`` over NCCL: a broadcast of each pushed vertex from rank 0 and an
all-gather of the per-epigraph partials, summed in fixed rank order on every rank.

This module is the host-side plumbing: it ships the ``ncclUniqueId`` through
``torch.distributed`` (any backend), mirrors the library's partition arithmetic
(``csrc/common.cuh``: owner_of / local_of / local_count) and reassembles sharded per-scenario
results into global order.
"""
from __future__ import annotations

import numpy as np

TILE = 128


def owner_of(g, world: int):
    return (np.asarray(g) // TILE) % world


def local_of(g, world: int):
    g = np.asarray(g)
    return (g // (TILE * world)) * TILE + g % TILE


def local_count(n_global: int, rank: int, world: int) -> int:
    full, rem = divmod(int(n_global), TILE)
    n = (full // world + (1 if full % world > rank else 0)) * TILE
    if rem and full % world == rank:
        n += rem
    return n


def global_of(rank: int, local, world: int):
    local = np.asarray(local)
    return ((local // TILE) * world + rank) * TILE + local % TILE


def owned_ordinals(n_global: int, rank: int, world: int) -> np.ndarray:
    """Global ordinals held by ``rank``, in its local order."""
    return global_of(rank, np.arange(local_count(n_global, rank, world)), world)


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 calls ``make_id()`` (-> 128 bytes); everyone returns the same bytes."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) != 128:
        raise RuntimeError("bad ncclUniqueId received")
    return bytes(box[0])


def init_context(device: int | None = None, group=None):
    """Create the library context of this rank of an initialised ``torch.distributed`` job."""
    import os
    import torch.distributed as dist
    from . import twosd
    if not dist.is_initialized():
        return twosd.Context(device or 0)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    if world == 1:
        return twosd.Context(device)
    uid = exchange_unique_id(twosd.Context.nccl_unique_id, group)
    return twosd.Context(device, rank, world, uid)


def gather_scenario_results(local_values, n_global: int, group=None):
    """All-gather a per-local-scenario array and return it in global scenario order."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local_values = np.ascontiguousarray(local_values)
    assert len(local_values) == local_count(n_global, rank, world)
    cap = max(local_count(n_global, r, world) for r in range(world))
    buf = torch.zeros(cap, dtype=torch.from_numpy(local_values[:0].copy()).dtype)
    buf[:len(local_values)] = torch.from_numpy(local_values)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.zeros(n_global, dtype=local_values.dtype)
    for r in range(world):
        idx = owned_ordinals(n_global, r, world)
        out[idx] = parts[r].numpy()[:len(idx)]
    return out


def ordered_rank_sum(local_partial, group=None):
    """Host twin of the library's k_rank_sum: all-gather, then sum in rank order 0..world-1
    so every rank gets the same bits."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_partial, dtype=np.float64))
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    acc = np.zeros_like(parts[0].numpy())
    for r in range(world):
        acc = acc + parts[r].numpy()
    return acc
