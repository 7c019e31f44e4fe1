#!/bin/bash
# Fourth GPU call of this session: GPU suite after (a) the dense builder's key space = cells of the map, (b) vr_cube_grow reading
# 16 entries per thread; launch list of one bench run (how long the directed-grid build takes now); C3 bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c_gputests4.log
tail -3 gpurun_out/r2c_gputests4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches4.csv python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_bench4.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/r2c_launches4.csv')) if len(r)>5 and r[0].isdigit()]
d=defaultdict(lambda:[0,0.0])
for r in rows:
    try: v=float(r[-1].replace(',',''))
    except: continue
    d[r[4]][0]+=1; d[r[4]][1]+=v
for k,(n,t) in sorted(d.items(), key=lambda x:-x[1][1])[:8]: print(n, round(t/1e3,1), 'us', k[:80])
PY
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench4_c3.json 2> gpurun_out/r2c_bench4_c3.err
python -c "
import json
j=json.loads(open('gpurun_out/r2c_bench4_c3.json').read().strip().splitlines()[-1]); print('c3 ms', j['ms_per_step'], 'frac', j['roofline']['frac'], j['config'].get('octree_build'))"
