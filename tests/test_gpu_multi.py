"""Multi-GPU tests (need >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
One process per GPU over NCCL; the two sharding modes of freesasa_b200/parallel.py driven by the real
engine and compared with the CPU oracle.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    import freesasa_b200 as fs
    from freesasa_b200 import parallel, workloads
    from oracle import bindings as ob

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = fs.Engine(rank)
    failed = []

    def check(name, cond):
        if not bool(cond):
            failed.append(name)

    # ---- C4 shape: batch of independent structures, LPT-sharded, one all-gather
    structs = workloads.batch(9, 300, 900, seed=2)
    sizes = [len(r) for _, r in structs]

    def compute_mine(indices):
        if not indices:
            return torch.zeros(0, dtype=torch.float64, device=dev)
        outs = eng.calc_batch(fs.LEE_RICHARDS, [structs[k] for k in indices], 1.4, 50)
        return torch.from_numpy(np.concatenate(outs)).to(dev)

    outs = parallel.calc_batch_sharded(sizes, compute_mine)
    for k, (x, r) in enumerate(structs):
        err = np.abs(outs[k].cpu().numpy() - ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 50)).max()
        check(f"batch structure {k} vs oracle", err < 5e-4)
    # ---- C5 shape: one replicated structure, sorted-order output ranges, one all-gather, local un-permute
    x, r = workloads.capsid(40000, r_out=70.0, seed=3)
    dx, dr = torch.tensor(x, device=dev), torch.tensor(r, device=dev)

    def compute_shard(rk, w):
        return eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 40, shard=(rk, w))

    got = parallel.calc_replicated_sharded(len(r), compute_shard, eng.unpermute).cpu().numpy()
    whole = eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 40).cpu().numpy()
    check("replicated+sharded (sync) equals one pass", np.array_equal(got, whole))
    check("replicated+sharded vs oracle", np.abs(got - ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 40)).max() < 5e-4)
    # ---- the same exchange without a host round trip: kernels enqueued, all-gather queued behind them, ONE synchronisation
    n = len(r)
    bounds = parallel.shard_bounds(n, world)
    width = max(e - b for b, e in bounds)
    sorted_mine = torch.zeros(n, dtype=torch.float64, device=dev)
    padded = torch.zeros(width, dtype=torch.float64, device=dev)
    gathered = torch.empty(world * width, dtype=torch.float64, device=dev)

    def enqueue():
        return eng.calc_device_async(fs.LEE_RICHARDS, dx, dr, 1.4, 40, shard=(rank, world), out=sorted_mine)

    def gather(local):
        b, e = bounds[rank]
        padded[: e - b].copy_(local[b:e])
        dist.all_gather_into_tensor(gathered, padded)
        return gathered

    g = parallel.gather_after_enqueue(eng, enqueue, gather).view(world, width)
    full = torch.cat([g[q, : e - b] for q, (b, e) in enumerate(bounds)])
    check("two-half call + queued all-gather equals one pass", np.array_equal(eng.unpermute(full).cpu().numpy(), whole))
    # ---- and without any collective call: peer stores from the kernel epilogue + flag barriers over NVLink (CUDA IPC)
    pg = parallel.PeerGather(eng, n, rank, world)
    for it in range(3):
        got_peer = pg.step(lambda out: eng.calc_device_async(fs.LEE_RICHARDS, dx, dr, 1.4, 40, shard=(rank, world), out=out))
        check(f"peer-store all-gather, step {it}: max diff {np.abs(got_peer.cpu().numpy() - whole).max():.3g}",
              np.array_equal(got_peer.cpu().numpy(), whole))
    pg.close()
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if not failed else "FAILED: " + "; ".join(failed))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.timeout(600)
def test_two_ranks_nccl(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1"] * world


# ---- the C entry point: one process, several GPUs (include/fsb200.h: fsb200_calc_multi) --------------------------------
@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_calc_multi_one_structure_replicated_and_sharded():
    """C5 shape through fsb200_calc_multi(): inputs replicated over NVLink, outputs partitioned, areas peer-stored into
    device 0.  Bit-identical to the one-device call, and within tolerance of the oracle."""
    import freesasa_b200 as fs
    from oracle import bindings as ob

    x, r = fs.workloads.capsid(200000, r_out=120.0, seed=2)
    one = fs.calc_multi(fs.LEE_RICHARDS, [(x, r)], 1.4, 50, n_devices=1)[0]
    for n_dev in (2, _n_gpus()):
        got = fs.calc_multi(fs.LEE_RICHARDS, [(x, r)], 1.4, 50, n_devices=n_dev)[0]
        np.testing.assert_array_equal(got, one)
        st = fs.multi_stats()
        assert st["n_devices"] == n_dev and st["n_atoms"] == len(r) and all(ms > 0 for ms in st["integrate_ms"])
    assert np.abs(one - ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 50)).max() < 2e-4
    sr = fs.calc_multi(fs.SHRAKE_RUPLEY, [(x, r)], 1.4, 200, n_devices=2)[0]
    np.testing.assert_array_equal(sr, fs.calc_multi(fs.SHRAKE_RUPLEY, [(x, r)], 1.4, 200, n_devices=1)[0])
    # a structure too small to share falls back to one device; bad input fails loudly on every path
    xs, rs = fs.workloads.globule(3000, seed=1)
    np.testing.assert_array_equal(fs.calc_multi(0, [(xs, rs)], 1.4, 20, 2)[0], fs.calc_multi(0, [(xs, rs)], 1.4, 20, 1)[0])
    bad = x.copy()
    bad[12345, 1] = np.nan
    with pytest.raises(RuntimeError):
        fs.calc_multi(fs.LEE_RICHARDS, [(bad, r)], 1.4, 50, n_devices=2)
    with pytest.raises(RuntimeError):
        fs.calc_multi(fs.LEE_RICHARDS, [(x, r)], 1.4, 50, n_devices=64)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_calc_multi_batch_dealt_by_size():
    """C4 shape through fsb200_calc_multi(): structures dealt to the devices (LPT on atom counts); every structure equals
    the one-device result bit for bit."""
    import freesasa_b200 as fs

    structs = fs.workloads.batch(96, 2000, 6000, seed=4)
    one = fs.calc_multi(fs.LEE_RICHARDS, structs, 1.4, 50, n_devices=1)
    two = fs.calc_multi(fs.LEE_RICHARDS, structs, 1.4, 50, n_devices=_n_gpus())
    for a, b in zip(one, two):
        np.testing.assert_array_equal(a, b)
    assert fs.multi_stats()["n_structures"] == 96
