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
    ok = True
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
        ok &= bool(err < 5e-4)
    # ---- C5 shape: one replicated structure, sorted-order output ranges, one all-gather, local un-permute
    x, r = workloads.capsid(40000, r_out=70.0, seed=3)
    dx, dr = torch.tensor(x, device=dev), torch.tensor(r, device=dev)

    def compute_shard(rk, w):
        return eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 40, shard=(rk, w))

    got = parallel.calc_replicated_sharded(len(r), compute_shard, eng.unpermute).cpu().numpy()
    whole = eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 40).cpu().numpy()
    ok &= bool(np.array_equal(got, whole))
    ok &= bool(np.abs(got - ob.oracle_calc(x, r, ob.LEE_RICHARDS, 1.4, 40)).max() < 5e-4)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.timeout(600)
def test_two_ranks_nccl(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1"] * world
