#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: headline counters of every captured launch and the SASS lines that
collect the most warp-stall samples.  usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt [n_lines]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    n_lines = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    lines = ["ncu summary of %s (ncu --set full --clock-control none --import-source on)" % rep, ""]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        lines.append("launch: %s  grid %s" % (r[hdr.index("Kernel Name")][:100], r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""))
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                lines.append("    %-84s %-12s %s" % (h, units[i], r[i]))
        lines.append("")
    src = page(rep, "source")
    hdr, data = None, []
    for r in src:
        if r and r[0] == "Kernel Name":
            if data:
                break
            hdr = None
            continue
        if hdr is None:
            hdr = r
            continue
        data.append(r)
    if hdr and data:
        isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        cols = [c for c in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_math", "stall_mio",
                            "stall_lg", "stall_barrier", "stall_not_selected") if c in hdr]
        tot = sum(int(r[isamp]) for r in data)
        lines.append("warp-stall samples of the first launch: %d over %d SASS lines; top %d lines (index, SASS, samples, executed, %s)"
                     % (tot, len(data), n_lines, ", ".join(cols)))
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n_lines]):
            r = data[i]
            lines.append("  %5d  %-72s %6s %9s  %s" % (i, r[isrc].strip()[:72], r[isamp], r[iex],
                                                      " ".join(r[hdr.index(c)] for c in cols)))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
