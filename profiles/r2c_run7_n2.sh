#!/bin/bash
# Last GPU call of the session (two GPUs, ~1 minute): bench.py as the driver launches it at N = 2 and N = 1 after the change of the clock sampling.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c_check_c3_n2.json 2> gpurun_out/r2c_check_c3_n2.err
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_check_c3_n1.json 2> gpurun_out/r2c_check_c3_n1.err
python - <<'PY'
import json
for n in (2, 1):
    try:
        j=json.loads(open(f"gpurun_out/r2c_check_c3_n{n}.json").read().strip().splitlines()[-1])
        print("n", n, "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "per-rank", j["config"].get("per_rank_ms_per_frame"), "clocks", j["clocks"], "launches", j.get("gpu_launches"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -4 gpurun_out/r2c_check_c3_n2.err; tail -2 gpurun_out/r2c_check_c3_n1.err
