#!/bin/bash
# Final multi-GPU measurements of round 2 (gpurun --gpus 8 -- bash profiles/r2_final_scale.sh)
cd "$(dirname "$0")/.."
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_scale_c3_n1.json 2> gpurun_out/r2_final_scale_c3_n1.err
for n in 2 4 8; do run r2_final_scale_c3_n$n $n --steps 50 --warmup 5; done
run r2_final_scale_c3_2lights_n8 8 --steps 50 --warmup 5 --lights 2
run r2_final_scale_c4_n8 8 --steps 20 --warmup 5 --config c4
run r2_final_scale_c5_n8 8 --steps 10 --warmup 3 --config c5
run r2_final_scale_reference_n8 8 --steps 3 --warmup 3 --impl reference
python -m pytest tests/test_gpu_mgpu.py -q 2>&1 | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_final_scale_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print(f.split("/")[-1], "ms", round(j["ms_per_step"],4), "value", round(j["value"],1), "frac", r.get("frac"), "e2e ms", (j.get("e2e") or {}).get("ms_per_step"), "per-rank", j["config"].get("per_rank_ms_per_frame"), j["config"].get("frame_checksum"), j["config"].get("device_frame_checksum"))
    except Exception as e:
        print(f, "failed", e)
PY
