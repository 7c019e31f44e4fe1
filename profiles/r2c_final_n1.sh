#!/bin/bash
# Final single-GPU measurements of round 2, second session (gpurun -- bash profiles/r2c_final_n1.sh): GPU test suite,
# smoke(), bench lines of every config (default build: directed top grids, hit-block diet, solid-subtree collapse),
# reference arm, launch list of the default bench command under ncu, one ncu --set full capture of the ray kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -45 > gpurun_out/r2c_final_gputests.log
tail -3 gpurun_out/r2c_final_gputests.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2c_final_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_final_bench_c3_n1.json 2> gpurun_out/r2c_final_bench.err
python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/r2c_final_reference_c3_n1.json 2>> gpurun_out/r2c_final_bench.err
python bench.py --gpus 1 --steps 20 --warmup 5 --lights 2 --no-cpu-baseline > gpurun_out/r2c_final_bench_c3_2lights_n1.json 2>> gpurun_out/r2c_final_bench.err
for cfg in c1 c2 c4 c5; do python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_final_bench_${cfg}_n1.json 2>> gpurun_out/r2c_final_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_final_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_final_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vr_svo -s 4 -c 1 -f -o gpurun_out/r2c_svo_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_final_*n1.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print(f.split("/")[-1], "ms", round(j["ms_per_step"],4), "value", round(j["value"],1), "frac", r.get("frac"), "e2e", j["e2e"]["value"] if j.get("e2e") else None, "cpu", (j.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2c_final_bench.err
