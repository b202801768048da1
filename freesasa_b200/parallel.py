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


# ---------------------------------------------------------------------------------------------------------
# Without a host round trip between the kernel and the collective
# ---------------------------------------------------------------------------------------------------------
def gather_after_enqueue(engine, enqueue: Callable[[], "object"], gather: Callable[["object"], "object"]):
    """Two-half step: ``enqueue()`` puts the kernels on the stream (Engine.calc_device_async) and returns the device
    tensor they write; ``gather(tensor)`` queues the collective BEHIND them on the same stream; only then comes the one
    synchronisation (Engine.finish).  If the engine had to run its rare second pass (atoms with > 160 neighbours) after
    the collective was queued, the collective is issued again so that no rank keeps stale values."""
    local = enqueue()
    out = gather(local)
    if engine.finish() != 0:  # FSB200_SECOND_PASS
        out = gather(local)
        import torch

        torch.cuda.current_stream().synchronize()
    return out


class PeerGather:
    """The all-gather of per-atom areas WITHOUT a collective call: every rank owns symmetric output buffers (CUDA IPC),
    maps the buffers of all peers, and its integration kernels store each area into ALL of them (8 B per atom and peer over
    NVLink).  A step is
        calc_device_async (peer stores) -> done barrier -> finish
    where the barrier is a one-warp kernel flipping epoch flags in peer memory (fsb200_ctx_peer_barrier): no NCCL call and
    no host synchronisation between the kernels and the exchange.  Two buffers alternate from step to step, which is what
    makes a second ("everybody has consumed the previous result") barrier unnecessary: a rank can only be writing step
    k + 2 into the buffer of step k after every rank has passed the barrier of step k + 1, i.e. after every rank has begun
    step k + 1 and is therefore done with the result of step k.  torch.distributed is used ONCE, at construction, to
    exchange the 64-byte IPC handles."""

    def __init__(self, engine, n_total: int, rank: int, world: int, slot=None):
        """n_total doubles per symmetric buffer.  slot = (offset, count): this rank's kernel fills only that window of
        every buffer (independent structures, one per rank, gathered side by side); None: the kernel indexes the whole
        buffer (one replicated structure, outputs partitioned by shard)."""
        import torch
        import torch.distributed as dist

        from . import IpcBuffer

        self.engine, self.rank, self.world, self.n = engine, rank, world, int(n_total)
        self.slot = slot
        dev = engine.device
        self.mine = [IpcBuffer(dev, 8 * self.n) for _ in range(2)]
        self.flags = IpcBuffer(dev, 4 * 64)
        handles = [None] * world
        dist.all_gather_object(handles, (self.mine[0].handle, self.mine[1].handle, self.flags.handle))
        self.peers, self.peer_flags = [[], []], []
        for r, (h0, h1, h_flag) in enumerate(handles):
            if r == rank:
                self.peers[0].append(self.mine[0])
                self.peers[1].append(self.mine[1])
                self.peer_flags.append(self.flags)
            else:
                self.peers[0].append(IpcBuffer.open(dev, h0, 8 * self.n))
                self.peers[1].append(IpcBuffer.open(dev, h1, 8 * self.n))
                self.peer_flags.append(IpcBuffer.open(dev, h_flag, 4 * 64))
        self.outs = [m.tensor(torch.float64, self.n) for m in self.mine]
        self.windows = [o[slot[0]:slot[0] + slot[1]] if slot else o for o in self.outs]
        self.parity = 0
        self.out = self.outs[0]
        # nine areas in ten are exactly 0 (buried atoms): they are not stored remotely; every rank zeroes the buffer of the
        # NEXT step ahead of its barrier signal of this step instead (stream order makes the zeroing visible before any peer
        # can pass that barrier and start writing into it)
        for o in self.outs:
            o.zero_()
        torch.cuda.current_stream().synchronize()
        engine.set_peer_zero_skipping(True)
        dist.barrier()

    def barrier(self):
        self.engine.peer_barrier(self.rank, self.world, [f.ptr for f in self.peer_flags])

    def step(self, enqueue: Callable[["object"], "object"]):
        """enqueue(out) must call engine.calc_device_async(..., out=out).  Returns the complete result (this rank's
        symmetric buffer of this step; valid until the step after the next) after the one synchronisation."""
        k = self.parity
        self.parity ^= 1
        off = 8 * int(self.slot[0]) if self.slot else 0
        self.engine.set_peer_outputs([p.ptr + off for r, p in enumerate(self.peers[k]) if r != self.rank])
        self.outs[k ^ 1].zero_()   # the buffer of the next step (its last result has been consumed: step() was called again)
        enqueue(self.windows[k])
        self.barrier()          # every rank's kernels — and with them every peer store into my buffer — are complete
        rc = self.engine.finish()
        if rc != 0:             # second pass ran after the barrier: agree on completion once more
            import torch

            self.barrier()
            torch.cuda.current_stream().synchronize()
        self.engine.peer_barrier_status()
        self.out = self.outs[k]
        return self.out

    def close(self):
        self.engine.set_peer_outputs([])
        self.engine.set_peer_zero_skipping(False)
        for k in range(2):
            for r, p in enumerate(self.peers[k]):
                if r != self.rank:
                    p.close()
        for r, f in enumerate(self.peer_flags):
            if r != self.rank:
                f.close()
