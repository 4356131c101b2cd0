set -x
nvidia-smi topo -m | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/bench8.err | tail -1 > gpurun_out/r2_bench_8gpu.json
tail -3 gpurun_out/bench8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_8gpu.json"))
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "with_nccl_gather", "strong_scaling")}, indent=1))
PY
