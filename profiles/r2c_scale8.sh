#!/bin/bash
# Eight GPUs with the directed top grids (gpurun --gpus 8 -- bash profiles/r2c_scale8.sh): C3 at 1 / 4 / 8 GPUs (50 frames) and at 8 with the
# contract's 20 frames, C4 at 8.  Kept short: an 8-GPU box is charged eight times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
}
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_scale8_c3_n1.json 2> gpurun_out/r2c_scale8_c3_n1.err &
wait
run r2c_scale8_c3_n8 8 --steps 50 --warmup 5
run r2c_scale8_c3_n4 4 --steps 50 --warmup 5
run r2c_scale8_c3_n8_steps20 8 --steps 20 --warmup 5
run r2c_scale8_c4_n8 8 --steps 20 --warmup 5 --config c4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_scale8_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms", round(j["ms_per_step"],4), "value", round(j["value"],1), "e2e ms", (j.get("e2e") or {}).get("ms_per_step"), "per-rank", j["config"].get("per_rank_ms_per_frame"), j["config"].get("frame_checksum"), j["config"].get("device_frame_checksum"), "clock samples", j["clocks"]["samples"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/r2c_scale8_c3_n8.err
