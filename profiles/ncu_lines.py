"""Per-source-line instruction counts of an ncu capture (--set full --import-source on): which lines of the device code
the warp instructions of a kernel were issued for.  Runs here on the CPU box.

    python profiles/ncu_lines.py gpurun_out/x.ncu-rep [top=60]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, agg, tot = None, None, [], 0
sass_tot = 0
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_exec = hdr.index("Instructions Executed")
        continue
    if not r or hdr is None:
        continue
    if r[0].isdigit() and r[i_exec].isdigit():
        agg.append((cur, int(r[0]), int(r[i_exec]), r[1]))
    elif r[0] == "" and len(r) > i_exec and r[i_exec].isdigit():
        sass_tot += int(r[i_exec])
tot = sum(a[2] for a in agg)
print(f"warp instructions attributed to source lines: {tot / 1e6:.1f} M (SASS total {sass_tot / 1e6:.1f} M; inlined code is attributed once per inlining level)")
agg.sort(key=lambda a: -a[2])
cum = 0
for f, l, e, s in agg[:top]:
    cum += e
    print(f"{f}:{l:4d} {e / 1e6:8.1f}M {100 * e / tot:5.2f}% cum {100 * cum / tot:5.1f}%  {s.strip()[:120]}")
