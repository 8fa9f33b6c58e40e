"""Summarise ncu captures into small text files under profiles/ (the .ncu-rep files stay in gpurun_out/).

    python tools/summarize_ncu.py report gpurun_out/k_score_r01.ncu-rep profiles/r01_k_score_full.txt
    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def report(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = ["# %s  (ncu --set full --clock-control none; one launch)" % rep]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("kernel: %s" % d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                lines.append("  %-86s %s %s" % (k, d[k], units[hdr.index(k)]))
    open(out, "w").write("\n".join(lines) + "\n")


def launches(csvf, out):
    rows = [r for r in csv.reader(open(csvf)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    lines = ["# %s  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)" % csvf,
             "%-60s %8s %14s %8s" % ("kernel", "launches", "total_us", "share")]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-60s %8d %14.1f %7.1f%%" % (k[:60], n, ns / 1e3, 100 * ns / tot))
    # the timed step proper = the launches over the full batch (grid has a 1024 in it); bench.py's later sections
    # (single-frame latency loop, training rounds) launch the same kernels on tiny grids
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    if gi is not None:
        big = OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= vi or r[mi] != "gpu__time_duration.sum" or "1024" not in r[gi] or "dsac::" not in r[ki]:
                continue
            ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1)
            a = big.setdefault(r[ki].split("(")[0], [0, 0.0])
            a[0] += 1
            a[1] += ns
        if big:
            n_steps = min(a[0] for a in big.values())
            tot_b = sum(a[1] / a[0] for a in big.values())
            lines += ["", "# full-batch launches only (1024 frames): average per launch and share of one step",
                      "%-60s %8s %14s %8s" % ("kernel", "launches", "avg_us", "share")]
            for k, (n, ns) in sorted(big.items(), key=lambda kv: -kv[1][1] / kv[1][0]):
                lines.append("%-60s %8d %14.1f %7.1f%%" % (k[:60], n, ns / n / 1e3, 100 * (ns / n) / tot_b))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
