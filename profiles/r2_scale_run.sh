python -m pytest tests/test_gpu_mgpu.py -x -q 2>&1 | tail -4
python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 100 --warmup 5 --gather direct > gpurun_out/r2_scale_n8_legacy_direct.json 2> gpurun_out/r2_scale_n8_legacy.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --impl reference > gpurun_out/r2_ref_n8.json 2> gpurun_out/r2_ref_n8.err
python - <<'PY'
import json
for n in ("n1","n2","n4","n8","n8_legacy_direct"):
    try:
        j=json.loads(open(f"gpurun_out/r2_scale_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms", round(j["ms_per_step"],4), "Mrays/s", round(j["value"]), "e2e ms", round(j["e2e"]["ms_per_step"],4), "e2e Mrays/s", round(j["e2e"]["value"]), "chk", j["config"]["frame_checksum"], j["config"].get("device_frame_checksum"))
    except Exception as e:
        print(n, "failed", e)
j=json.loads(open("gpurun_out/r2_ref_n8.json").read().strip().splitlines()[-1]); print("ref n8", j["value"], j["cpu_baseline"])
PY
tail -3 gpurun_out/r2_scale_n8.err
