#!/bin/bash
# Round 2, second session, first GPU call: the directed top grids (eight per-octant tables) + the exact hit-block diet.
# GPU suite, A/B on one box (main = directed grids + diet; bench.py also times the undirected grid in the same process;
# lit = -DVR_HIT_LITERAL: the general IEEE divisions; grid21 = coarser grid), C4 / C2, one ncu --set full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2b_gputests.log
tail -3 gpurun_out/r2b_gputests.log
bash profiles/ab.sh run c3 main lit grid21 main 2>&1 | tee gpurun_out/r2b_ab1.txt
bash profiles/ab.sh run c4 main 2>&1 | tee -a gpurun_out/r2b_ab1.txt
bash profiles/ab.sh run c2 main 2>&1 | tee -a gpurun_out/r2b_ab1.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vr_svo -s 4 -c 1 -f -o gpurun_out/r2b_svo_directed python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out
