#!/usr/bin/env python
"""bench.py — headline benchmark of the SASA hot path (BASELINE.json: atoms/sec, Lee-Richards
n_slices=100, 100k-atom synthetic globule; max |dSASA| vs the reference).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

A "step" is one pass of the hot path (cell list build + Lee-Richards integration of every atom) over
one structure per GPU.  N=1: config C2.  N>1 (launched by torch.distributed.run, one rank per GPU):
every rank integrates its own 100k-atom structure (weak scaling: work per GPU fixed) and the step ends
with ONE NCCL all-gather of the per-atom SASA of all ranks; value = atoms of all ranks / max-over-ranks time.

Reported numbers
  value    atoms/s with the inputs already resident in HBM; device time from CUDA events on the stream
           the kernels are launched on; L2 flushed (256 MiB write) between timed steps.
  e2e      atoms/s through the reference-facing call freesasa_calc_coord() of the C host layer with
           HOST buffers (N=1), i.e. H2D of xyz+radii and D2H of the areas inside the timed region
           (N>1: pinned host -> device -> compute -> all-gather -> host on every rank).
  roofline algorithmic bytes (40 B/atom: 24 B xyz + 8 B radius in, 8 B area out) / integrate-kernel
           time vs the measured HBM copy bandwidth.  The integration kernel is FP32-issue bound, not
           HBM bound (~5.5e3 circle-circle evaluations per atom), so this fraction is tiny by
           construction; `compute` adds pair-slice evaluations/s.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on this box's host cores, same arrays.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ATOMS = 100_000
N_SLICES = 100
PROBE = 1.4
ALG_BYTES_PER_ATOM = 40.0
# DRAM traffic of the dominant kernel, from the committed ncu --set full capture of this same command
# (bench.py --steps 2 --warmup 3): 6.00 MB read + 0.0 KB written per launch for 100k atoms = 60 B/atom,
# i.e. the 32 B/atom sorted double4 records are fetched about once (plus the item queue, the permutation and the L2-flushed
# output lines) and everything else stays in L2/shared memory.
NCU_DRAM_BYTES_PER_LAUNCH = 6001920
NCU_SOURCE = "profiles/r1_15_lr_cert_antipodal_hybrid.txt"
METRIC = "atoms/sec (LR n_slices=100)"
WORKLOAD = "C2: 100k-atom synthetic globular coord array, Lee-Richards n_slices=100, probe 1.4 A"


def host_threads():
    return max(1, min(16, os.cpu_count() or 1))  # 16 = the reference's hard cap (src/sasa_lr.c:17)


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
            "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
            "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
            "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_run(xyz, radii, threads):
    """One pass of the reference's own CPU implementation (oracle/_ref when built, else our C port)."""
    from oracle import bindings as ob

    if ob.ref_available():
        ob.ref_lib().freesasa_set_verbosity(1)
        t = time.perf_counter()
        sasa = ob.ref_calc(xyz, radii, ob.LEE_RICHARDS, PROBE, N_SLICES, threads)
        return time.perf_counter() - t, sasa, "reference"
    t = time.perf_counter()
    sasa = ob.oracle_calc(xyz, radii, ob.LEE_RICHARDS, PROBE, N_SLICES, threads)
    return time.perf_counter() - t, sasa, "port"


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from freesasa_b200 import workloads

    xyz, radii = workloads.globule(N_ATOMS, seed=0)
    threads = host_threads()
    for _ in range(args.warmup):
        cpu_reference_run(xyz, radii, threads)
    times, kind = [], "reference"
    for _ in range(args.steps):
        dt, _, kind = cpu_reference_run(xyz, radii, threads)
        times.append(dt)
    total = float(np.sum(times))
    value = N_ATOMS * args.steps / total
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "atoms": N_ATOMS, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": threads, "kind": kind,
                         "sample": "full 100k-atom structure per step, freesasa_calc_coord wall clock"},
        "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner to fd 1)
    must not pollute it: point fd 1 at stderr for the rest of the run and keep the real stdout aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-certificate", action="store_true",
                    help="ablation: integrate every atom, do not use the buried-atom certificate (results are identical)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    import freesasa_b200 as fs
    from freesasa_b200 import workloads

    if not fs.available():
        raise SystemExit("bench.py: the CUDA engine is not built or no B200 is visible (there is no CPU fallback)")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: one structure per rank ------------------------------------------------------------
    xyz, radii = workloads.globule(N_ATOMS, seed=rank)
    h_xyz = torch.from_numpy(xyz).pin_memory()
    h_rad = torch.from_numpy(radii).pin_memory()
    d_xyz, d_rad = h_xyz.to(dev), h_rad.to(dev)
    d_out = torch.zeros(N_ATOMS, dtype=torch.float64, device=dev)
    d_all = torch.zeros(world * N_ATOMS, dtype=torch.float64, device=dev) if distributed else None
    h_all = torch.empty(world * N_ATOMS, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    align = torch.zeros(1, device=dev)
    eng = fs.Engine(local_rank, fs.FP32)
    if args.no_certificate:
        eng.set_certificate(False)
    stream = torch.cuda.Stream(dev)  # a real (non-legacy) stream: the engine replays its launch sequence as a CUDA graph on it
    torch.cuda.set_stream(stream)

    def device_step():
        eng.calc_device(fs.LEE_RICHARDS, d_xyz, d_rad, PROBE, N_SLICES, out=d_out)
        if distributed:
            dist.all_gather_into_tensor(d_all, d_out)

    def e2e_step():
        if not distributed:
            p = fs.Parameters(fs.LEE_RICHARDS, PROBE, 100, N_SLICES, 1)
            return fs.calc_coord(xyz, radii, p).sasa  # host arrays in, host array out (the drop-in call)
        d_xyz.copy_(h_xyz, non_blocking=True)
        d_rad.copy_(h_rad, non_blocking=True)
        device_step()
        h_all.copy_(d_all, non_blocking=True)
        torch.cuda.synchronize(dev)
        return h_all

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up ---------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    e2e_step()
    barrier()

    # ---- timed region: device-resident ---------------------------------------------------------------
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    integrate_ms, device_ms = [], []
    launches0 = fs.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)  # evict L2 between timed steps (outside the timed events)
            if distributed:
                dist.all_reduce(align)  # ranks leave the (untimed) flush together: the step's all-gather then
                # only waits for differences in compute time, not for flush/launch skew between ranks
            ev0[k].record(stream)
            device_step()
            ev1[k].record(stream)
            st = eng.stats()
            certified = st["n_certified"]
            integrate_ms.append(st["integrate_ms"])
            device_ms.append(st["device_ms"])
        barrier()
    launches = fs.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * N_ATOMS * args.steps / (total_ms * 1e-3)

    # ---- timed region: end to end with host buffers ------------------------------------------------------
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * N_ATOMS * e2e_steps / float(e2e_s.item())

    if rank == 0:
        got = np.asarray(res[:N_ATOMS]) if distributed else res
        peak, peak_src = measured_peak_hbm()
        k_ms = float(np.mean(integrate_ms))
        achieved = ALG_BYTES_PER_ATOM * N_ATOMS / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "atoms_per_gpu": N_ATOMS, "structures": world,
                       "buried_atom_certificate": ("off (ablation)" if args.no_certificate else
                                                   f"on: {certified} of {N_ATOMS} atoms proved fully buried, not integrated (exact, DESIGN.md)"),
                       "l2": "flushed between timed steps (256 MiB device write, outside the events)",
                       "timing": "CUDA events on the launching stream, sum over steps, max over ranks",
                       "collective": "one NCCL all-gather of per-atom SASA per step" if distributed else "none (N=1)"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "atoms/s", "h2d_bytes_per_step": 32 * N_ATOMS * world,
                    "d2h_bytes_per_step": 8 * N_ATOMS * world * (world if distributed else 1), "steps": e2e_steps,
                    "path": "freesasa_calc_coord() of the C host layer, pageable host arrays" if not distributed
                    else "pinned host -> H2D -> calc_device -> NCCL all-gather -> D2H on every rank"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_integrate<LR,float>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, " + NCU_SOURCE + ")",
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_ATOM * N_ATOMS, "peak_source": peak_src,
                         "algorithmic_bytes_per_atom": ALG_BYTES_PER_ATOM, "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / (total_ms / args.steps),
                         "note": "FP32-issue bound kernel; HBM fraction is tiny by construction (DESIGN.md)"},
            "device_ms_per_call": float(np.mean(device_ms)),
        }
        if not args.no_cpu_baseline:
            threads = host_threads()
            cpu_s, want, kind = cpu_reference_run(xyz, radii, threads)
            err = np.abs(got - want)
            line["cpu_baseline"] = {"value": N_ATOMS / cpu_s, "unit": "atoms/s", "cores": threads, "kind": kind,
                                    "sample": "the full 100k-atom structure of rank 0, one pass, wall clock around freesasa_calc_coord"}
            line["parity"] = {"max_abs_dsasa": float(err.max()), "worst_atom": int(err.argmax()), "tolerance": 1e-3,
                              "total_gpu": float(got.sum()), "total_ref": float(want.sum())}
            nn_pairs = None
            try:
                from oracle import bindings as ob

                start, _ = ob.oracle_neighbours(xyz, radii + PROBE)
                nn_pairs = int(start[-1])
            except Exception:
                pass
            if nn_pairs:
                line["roofline"]["compute"] = {"pair_slice_evals_per_s": nn_pairs * N_SLICES / (k_ms * 1e-3),
                                               "pair_slice_evals": nn_pairs * N_SLICES}
        emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
