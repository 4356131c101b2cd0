set -x
nvidia-smi topo -m | head -8
python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2.err | tail -1 > gpurun_out/r2_bench_2gpu.json
tail -5 gpurun_out/bench2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "with_nccl_gather", "strong_scaling")}, indent=1))
PY
