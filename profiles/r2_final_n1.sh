#!/bin/bash
# Final single-GPU measurements of round 2 (gpurun -- bash profiles/r2_final_n1.sh): GPU test suite, bench lines of every config,
# launch list of the default bench command under ncu (kernel share of the step), smoke().
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_final_gputests.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2_final_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_c3_n1.json 2> gpurun_out/r2_final_bench_c3_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/r2_final_reference_c3_n1.json 2>> gpurun_out/r2_final_bench_c3_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 --lights 2 > gpurun_out/r2_final_bench_c3_2lights_n1.json 2>> gpurun_out/r2_final_bench_c3_n1.err
for cfg in c1 c2 c4 c5; do python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r2_final_bench_${cfg}_n1.json 2>> gpurun_out/r2_final_bench_c3_n1.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_ncu_bench.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_final_*n1.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print(f.split("/")[-1], "ms", round(j["ms_per_step"],4), "value", round(j["value"],1), "frac", r.get("frac"), "e2e", j["e2e"]["value"] if j.get("e2e") else None, "cpu", (j.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2_final_bench_c3_n1.err
