#!/bin/bash
# Second GPU call of this session: the whole GPU suite with the solid-subtree collapse (device post-process, all walks),
# then the bench line of C3 (is the ray kernel's time unchanged by the solid-node test in the descent loop?).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2b_gputests2.log
tail -4 gpurun_out/r2b_gputests2.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench2_c3.json 2> gpurun_out/r2b_bench2_c3.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2b_bench2_c3.json").read().strip().splitlines()[-1])
print("c3 ms", j["ms_per_step"], "frac", j["roofline"]["frac"], "other", j["config"].get("other_walk_ms_per_frame"), "build", j["config"].get("octree_build"))
PY
