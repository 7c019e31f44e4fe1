python -m pytest tests/test_gpu_mgpu.py tests/test_facade.py -q -m gpu 2>&1 | tail -2
for mode in "" "VR_MGPU_SIGNAL_KERNEL=1"; do
  env $mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_sig.json 2> gpurun_out/r2_sig.err
  python -c "
import json
j=json.loads(open('gpurun_out/r2_sig.json').read().strip().splitlines()[-1]); print('[$mode]', 'ms', round(j['ms_per_step'],4), 'value', round(j['value']), 'launches', j['gpu_launches'], 'e2e ms', round(j['e2e']['ms_per_step'],4), j['config']['frame_checksum'], j['config']['device_frame_checksum'])"
done
env python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29901 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_sig20.json 2> gpurun_out/r2_sig.err
python -c "
import json
j=json.loads(open('gpurun_out/r2_sig20.json').read().strip().splitlines()[-1]); print('steps 20:', 'ms', round(j['ms_per_step'],4), 'value', round(j['value']))"
tail -2 gpurun_out/r2_sig.err
