"""Multi-GPU orchestration: one process per GPU, torch.distributed for the plumbing.

The hot path has NO exchange step during compute — the area of an atom depends only on atoms within
R_i + R_max of it — so the work shards two ways and each ends with exactly one all-gather of per-atom SASA:

* ``calc_batch_sharded``        independent structures (config C4): structures are dealt to ranks by
                                longest-processing-time on their atom counts; every rank integrates its
                                structures in one batched device pass; one all-gather returns all areas
                                to all ranks.
* ``calc_replicated_sharded``   one huge structure (config C5): inputs replicated on every rank, every
                                rank builds the (cheap, deterministic) cell list and integrates only its
                                contiguous range of the cell-sorted atom order; one all-gather of the sorted
                                areas, then a local un-permute.  "Chain-sharded" in the reference's sense of
                                computing chains in isolation (src/structure.c:955-1081) is a different
                                quantity and is NOT what this does: every atom sees all neighbours.

Both take the compute step as a callable so the CPU (gloo, world_size 2) tests can exercise the
partition / gather / un-pad logic without a GPU; on the GPU box the callables are Engine methods.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def lpt_assign(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to `world` bins; deterministic."""
    order = sorted(range(len(sizes)), key=lambda k: (-int(sizes[k]), k))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda q: (load[q], q))
        bins[r].append(k)
        load[r] += int(sizes[k])
    for b in bins:
        b.sort()
    return bins


def shard_bounds(n: int, world: int) -> List[tuple]:
    """[begin, end) of every rank's share of n sorted atoms — same formula as fsb200_shard_begin/end."""
    return [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]


def _all_gather_padded(local, width: int):
    """all_gather of equal-width 1-D float64 tensors (local is padded to `width`)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    buf = torch.zeros(width, dtype=torch.float64, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * width, dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return out.view(world, width)


def calc_batch_sharded(sizes: Sequence[int], compute_mine: Callable[[List[int]], "object"], device=None):
    """sizes[k] = atoms of structure k (known to all ranks).  compute_mine(indices) returns a 1-D
    float64 tensor: the areas of this rank's structures concatenated in `indices` order.
    Returns a list of per-structure tensors (all structures, on every rank)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    bins = lpt_assign(sizes, world)
    local = compute_mine(bins[rank])
    width = max(1, max(sum(int(sizes[k]) for k in b) for b in bins))
    assert local.shape[0] == sum(int(sizes[k]) for k in bins[rank])
    gathered = _all_gather_padded(local, width)  # the ONE collective
    out = [None] * len(sizes)
    for r, b in enumerate(bins):
        off = 0
        for k in b:
            out[k] = gathered[r, off : off + int(sizes[k])]
            off += int(sizes[k])
    return out


def calc_replicated_sharded(n: int, compute_shard: Callable[[int, int], "object"], unpermute: Callable[["object"], "object"]):
    """compute_shard(rank, world) returns a length-n float64 tensor whose [begin,end) slice (this rank's
    share of the SORTED order) is filled.  unpermute(sorted) maps the gathered sorted areas back to the
    caller's atom order.  Returns the length-n tensor in caller order (on every rank)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(n, world)
    mine = compute_shard(rank, world)
    b, e = bounds[rank]
    width = max(1, max(hi - lo for lo, hi in bounds))
    gathered = _all_gather_padded(mine[b:e], width)  # the ONE collective
    import torch

    full = torch.cat([gathered[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)])
    return unpermute(full)
