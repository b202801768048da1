"""CPU tests of the multi-rank orchestration (world_size 2, gloo): partition, the single all-gather,
un-padding and order restoration.  The compute step is played by the CPU oracle here; on the GPU box
the same functions are driven by Engine (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from freesasa_b200 import parallel, workloads


def test_lpt_assign_balances_and_covers():
    sizes = [5000, 4100, 5900, 4800, 4000, 6000, 5100, 4500, 4700]
    for world in (1, 2, 4, 8):
        bins = parallel.lpt_assign(sizes, world)
        flat = sorted(k for b in bins for k in b)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[k] for k in b) for b in bins]
        assert max(loads) - min(loads) <= max(sizes)


def test_shard_bounds_partition():
    for n in (1, 10, 1000003):
        for w in (1, 2, 3, 8):
            b = parallel.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings as ob

    ob.oracle_lib().oracle_set_threads(2)
    # ---- batch of independent structures
    structs = workloads.batch(5, 150, 400, seed=3)
    sizes = [len(r) for _, r in structs]

    def compute_mine(indices):
        parts = [ob.oracle_calc(structs[k][0], structs[k][1], ob.LEE_RICHARDS, 1.4, 10) for k in indices]
        return torch.from_numpy(np.concatenate(parts)) if parts else torch.zeros(0, dtype=torch.float64)

    outs = parallel.calc_batch_sharded(sizes, compute_mine)
    ok = True
    for k, (x, r) in enumerate(structs):
        ok &= np.array_equal(outs[k].numpy(), ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 10))
    # ---- one replicated structure, output range sharded in a "sorted" order (here: sorted by x)
    x, r = workloads.globule(700, seed=5)
    perm = np.argsort(x[:, 0], kind="stable")
    full = ob.oracle_calc(x, r, ob.SHRAKE_RUPLEY, 1.4, 50)

    def compute_shard(rk, w):
        b, e = parallel.shard_bounds(len(r), w)[rk]
        out = torch.zeros(len(r), dtype=torch.float64)
        out[b:e] = torch.from_numpy(full[perm][b:e])  # this rank "integrates" only its sorted range
        return out

    def unpermute(sorted_vals):
        out = torch.empty_like(sorted_vals)
        out[torch.from_numpy(perm)] = sorted_vals
        return out

    got = parallel.calc_replicated_sharded(len(r), compute_shard, unpermute)
    ok &= np.array_equal(got.numpy(), full)

    # ---- the two-half step: kernels enqueued, collective queued behind them, ONE synchronisation; when the engine reports
    # a second pass (results completed after the collective had been queued) the collective must be issued again
    class FakeEngine:
        """Plays fsb200_ctx_calc_device_async / fsb200_ctx_finish: the first `late` shards are only complete after finish()."""

        def __init__(self, late):
            self.late, self.calls = late, 0

        def enqueue(self):
            b, e = parallel.shard_bounds(len(r), world)[rank]
            self.buf = torch.zeros(len(r), dtype=torch.float64)
            self.buf[b:e] = torch.from_numpy(full[perm][b:e])
            if self.late:
                self.hidden = self.buf[b].clone()
                self.buf[b] = -1.0          # not there yet
            return self.buf

        def finish(self):
            if self.late:
                b, _ = parallel.shard_bounds(len(r), world)[rank]
                self.buf[b] = self.hidden   # the second pass completes it
                self.late = False
                return 1                     # FSB200_SECOND_PASS
            return 0

    for late in (False, True):
        eng = FakeEngine(late)
        bounds = parallel.shard_bounds(len(r), world)
        width = max(e - b for b, e in bounds)
        n_gathers = [0]

        def gather(local):
            n_gathers[0] += 1
            b, e = bounds[rank]
            padded = torch.zeros(width, dtype=torch.float64)
            padded[: e - b] = local[b:e]
            out = torch.empty(world * width, dtype=torch.float64)
            dist.all_gather_into_tensor(out, padded)
            return out

        import sys
        import types

        fake_cuda = None
        if late and not torch.cuda.is_available():   # gather_after_enqueue synchronises the current CUDA stream after a redo
            fake_cuda = torch.cuda.current_stream
            torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(synchronize=lambda: None)
        try:
            g = parallel.gather_after_enqueue(eng, eng.enqueue, gather).view(world, width)
        finally:
            if fake_cuda is not None:
                torch.cuda.current_stream = fake_cuda
        merged = torch.cat([g[q, : e - b] for q, (b, e) in enumerate(bounds)])
        ok &= np.array_equal(unpermute(merged).numpy(), full)
        ok &= n_gathers[0] == (2 if late else 1)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]
