"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summary that is
committed under profiles/:  python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/NAME.txt "title"
Needs the ncu CLI (present in the dev container; no GPU required to READ a report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"


def ncu(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", f"# source report: {rep} (ncu --set full --clock-control none --import-source on)", ""]
    for vals in rows[2:]:
        rec = dict(zip(hdr, zip(units, vals)))
        lines.append("kernel: " + rec.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in rec:
                lines.append(f"  {k:75s} {rec[k][1]:>18s} {rec[k][0]}")
        lines.append("  warp stall reasons (warps stalled per issue-active cycle):")
        st = sorted(((float(v[1]), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in rec.items()
                     if k.startswith(STALLS) and v[1] not in ("", "n/a")), reverse=True)
        for val, name in st[:8]:
            lines.append(f"    {name:28s} {val:8.3f}")
        lines.append("")
    try:
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
        h = src[2]
        i_samp, i_inst, i_thr = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        agg = []
        for r in src[3:]:
            if len(r) > i_thr and r[2] == "-" and r[0].isdigit():
                try:
                    agg.append((int(r[i_inst]), int(r[i_samp]), int(r[0]), r[1].strip(), int(r[i_thr])))
                except ValueError:
                    pass
        ts = sum(a[1] for a in agg) or 1
        lines.append("top source lines by issued warp instructions (first kernel; inlined lines are counted where they are written):")
        for inst, samp, ln, text, thr in sorted(agg, reverse=True)[:25]:
            lines.append(f"  {inst:12d} inst  {100 * samp / ts:5.1f}% samples  {thr / max(inst, 1):4.1f} lanes  L{ln:<4d} {text[:100]}")
        lines.append("top source lines by stall samples:")
        for inst, samp, ln, text, thr in sorted(agg, key=lambda a: -a[1])[:10]:
            lines.append(f"  {100 * samp / ts:5.1f}% samples  {inst:12d} inst  L{ln:<4d} {text[:100]}")
    except Exception as e:  # source page is optional
        lines.append(f"(no source page: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
