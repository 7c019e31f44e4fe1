for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2_scale2_n$n.json 2> gpurun_out/r2_scale2_n$n.err
done
python - <<'PY'
import json
for n in ("n2","n4","n8"):
    try:
        j=json.loads(open(f"gpurun_out/r2_scale2_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms", round(j["ms_per_step"],4), "Mrays/s", round(j["value"]), "e2e ms", round(j["e2e"]["ms_per_step"],4), "e2e Mrays/s", round(j["e2e"]["value"]), "chk", j["config"]["frame_checksum"], j["config"].get("device_frame_checksum"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2_scale2_n8.err
