"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr = None, None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    if r[0] != "":            # a source line row: its own totals follow in the SASS rows
        key = (cur_file, r[0], r[1].strip()[:100])
        agg.setdefault(key, [0, 0])
        cur = key
        continue
    try:
        agg[cur][0] += int(r[ie].replace(",", ""))
        agg[cur][1] += int(r[ss].replace(",", ""))
    except Exception:
        pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (ti, ts))
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s:%s  %s" % (100.0 * i / ti, 100.0 * s / ts, f, ln, src))
