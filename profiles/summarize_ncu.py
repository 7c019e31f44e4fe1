"""Turns an ncu report (gpurun_out/*.ncu-rep, captured with --set full --import-source on) into the text
summary committed under profiles/.  Runs here on the CPU box: `ncu -i` needs no GPU.

    python profiles/summarize_ncu.py gpurun_out/prof_svo_r1.ncu-rep profiles/r1_svo_v1.txt "note ..."
"""
import csv
import io
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


lines = []
raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units = raw[0], raw[1]
for row in raw[2:]:
    name = row[hdr.index("Kernel Name")]
    lines.append(f"== kernel: {name}   grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}")
    for k in KEYS:
        if k in hdr:
            lines.append(f"   {k:75s} {row[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    stalls = [(float(row[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stalls:
        stalls = [(float(row[i] or 0), h) for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    lines.append("   -- top stall reasons")
    for v, h in sorted(stalls, reverse=True)[:6]:
        lines.append(f"   {h:75s} {v:18.3f}")

src = list(csv.reader(io.StringIO(ncu("source"))))
cols = src[1]
body = src[2:]
i_src, i_exec, i_samp = cols.index("Source"), cols.index("Instructions Executed"), cols.index("# Samples")
total = sum(int(r[i_exec]) for r in body if len(r) > i_exec and r[i_exec].isdigit())
samples = sum(int(r[i_samp]) for r in body if len(r) > i_samp and r[i_samp].isdigit())
lines.append(f"== SASS: {len(body)} instructions, {total} warp-instructions executed, {samples} stall samples")
mx = max(int(r[i_exec]) for r in body)
lines.append("   hottest loop (instructions executed > 50 % of the maximum), in program order:")
hot = [r for r in body if int(r[i_exec]) > 0.5 * mx]
lines.append(f"   {len(hot)} instructions, {sum(int(r[i_exec]) for r in hot)} executed = {100.0 * sum(int(r[i_exec]) for r in hot) / total:.1f} % of all, "
             f"{100.0 * sum(int(r[i_samp]) for r in hot) / max(samples, 1):.1f} % of samples")
for r in hot:
    lines.append(f"      {r[i_src].strip()[:64]:64s} exec {r[i_exec]:>12s} samples {r[i_samp]:>8s}")
mem = [r for r in body if any(op in r[i_src] for op in ("LDG", "STG", "LDS", "STS", "TEX", "TLD"))]
lines.append("   memory instructions:")
for r in mem:
    lines.append(f"      {r[i_src].strip()[:64]:64s} exec {r[i_exec]:>12s} samples {r[i_samp]:>8s}")

with open(out_path, "w") as f:
    if note:
        f.write(note + "\n")
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
