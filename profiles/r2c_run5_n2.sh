#!/bin/bash
# Two GPUs (gpurun --gpus 2): the library's multi-GPU scheduler with the directed top grids (every rank derives its own tables
# from the broadcast tree), the two-process GPU tests, the grid-builder tests after the one-pass (wavefront) cube builder,
# and the C3 bench line at N = 2 through torchrun as the driver launches it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_canonical.py tests/test_facade.py -q -m gpu -k "mgpu or top_grid or small or collapse or facade or random" 2>&1 | tail -6 | tee gpurun_out/r2c_gputests5_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2c_scale_c3_n2.json 2> gpurun_out/r2c_scale_c3_n2.err
python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_scale_c3_n1.json 2>> gpurun_out/r2c_scale_c3_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        j=json.loads(open(f"gpurun_out/r2c_scale_c3_n{n}.json").read().strip().splitlines()[-1])
        print("n", n, "ms", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["value"], "checksum", j["config"].get("frame_checksum"), j["config"].get("device_frame_checksum"), "clocks", j["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2c_scale_c3_n2.err
