#!/bin/bash
# Evidence for the technique choices north_star lists (run on ONE B200 box: gpurun -- profiles/r2_evidence.sh).
# For the octree kernel at the headline workload (c3): static 32x4 tiles vs persistent warps with warp-level refill
# (like for like: the same in-cell walk on both sides), and the L2 access-policy window over the nodes on / off --
# bench.py times (CUDA events, 30 frames) plus an ncu metrics pass of one frame (SIMT efficiency, issue-slot
# utilisation, instruction count, cache hit rates, DRAM bytes).  Writes gpurun_out/r2_persistent_l2_evidence.txt.
cd "$(dirname "$0")/.."
out=gpurun_out/r2_persistent_l2_evidence.txt
: > $out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
run() {   # name, bench flags
  name=$1; shift
  ms=$(python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(j['ms_per_step'],4), j['config']['frame_checksum'])")
  echo "== $name: bench.py $* -> ms/frame, checksum: $ms" | tee -a $out
  ncu --metrics $M --clock-control none -k regex:vr_svo -s 4 -c 1 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null \
    | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
for r in rows: print('   %-70s %s %s' % (r[-3], r[-1], r[-2]))" | tee -a $out
}
run "walk 1 (per-axis), static tiles"                     --walk 1
run "walk 1 (per-axis), persistent warps, refill at 8 idle lanes, 3 CTAs/SM"  --walk 1 --persistent 1 --refill-min 8 --ctas-per-sm 3
run "walk 1 (per-axis), persistent warps, refill at 24 idle lanes, 4 CTAs/SM" --walk 1 --persistent 1 --refill-min 24 --ctas-per-sm 4
run "walk 0 (merged), static tiles"                       --walk 0
run "walk 0 (merged), persistent warps, refill at 8 idle lanes, 3 CTAs/SM"    --walk 0 --persistent 1 --refill-min 8 --ctas-per-sm 3
run "walk 2 (closed form), static tiles"                  --walk 2
run "walk 2 (closed form), static tiles, L2 access-policy window over the nodes" --walk 2 --l2-persist 1
run "walk 1 (per-axis), static tiles, L2 access-policy window over the nodes"    --walk 1 --l2-persist 1
