"""Configs C4 and C5 of BASELINE.json on N GPUs (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
        tests/tools/measure_multi.py

C4: 1024 independent ~5k-atom structures, LR n_slices=50, structures dealt to ranks (LPT), one all-gather.
C5: one 1M-atom shell, LR n_slices=100, inputs replicated, each rank integrates its range of the sorted
    order, one all-gather, local un-permute.  Rank 0 writes gpurun_out/multi_N.json."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import freesasa_b200 as fs  # noqa: E402
from freesasa_b200 import parallel, workloads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = fs.Engine(local)
    out = {"n_gpus": world}

    def timed(fn, reps=3):
        fn()
        best = None
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize(dev)
            t = time.perf_counter()
            res = fn()
            torch.cuda.synchronize(dev); dist.barrier()
            dt = torch.tensor([time.perf_counter() - t], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = dt.item() if best is None else min(best, dt.item())
        return best, res

    # ---- C4 -----------------------------------------------------------------------------------------------
    structs = workloads.batch(1024, 4000, 6000, seed=0)
    sizes = [len(r) for _, r in structs]
    total = sum(sizes)

    def compute_mine(indices):
        outs = eng.calc_batch(fs.LEE_RICHARDS, [structs[k] for k in indices], 1.4, 50)
        return torch.from_numpy(np.concatenate(outs)).to(dev)

    t, res = timed(lambda: parallel.calc_batch_sharded(sizes, compute_mine))
    out["C4"] = {"structures": 1024, "atoms": total, "e2e_ms": t * 1e3, "atoms_per_s": total / t,
                 "note": "host arrays in on every rank's share, all per-atom areas on every rank's device out"}
    if rank == 0:
        from oracle import bindings as ob

        errs = [float(np.abs(res[k].cpu().numpy() - ob.oracle_calc(structs[k][0], structs[k][1], 0, 1.4, 50)).max()) for k in range(0, 1024, 128)]
        out["C4"]["max_err_sampled_8_structures"] = max(errs)

    # ---- C5 -----------------------------------------------------------------------------------------------
    x, r = workloads.capsid(1_000_000)
    dx, dr = torch.tensor(x, device=dev), torch.tensor(r, device=dev)

    def c5():
        return parallel.calc_replicated_sharded(len(r), lambda rk, w: eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 100, shard=(rk, w)), eng.unpermute)

    t, got = timed(c5)
    out["C5"] = {"atoms": len(r), "device_resident_ms": t * 1e3, "atoms_per_s": len(r) / t,
                 "note": "inputs replicated in HBM; each rank integrates its sorted range; one all-gather; local un-permute"}
    whole = eng.calc_device(fs.LEE_RICHARDS, dx, dr, 1.4, 100)
    out["C5"]["identical_to_single_gpu"] = bool(torch.equal(got, whole))
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"multi_{world}.json"), "w"), indent=1)
        sys.stderr.write(json.dumps(out, indent=1) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
