set -x
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1 --csv"
RBD_JIT_VARIANT=1 RBD_ONLY=smem RBD_NO_GATE=1 ncu $M -k regex:rbd_jit_smem -s 1 --log-file gpurun_out/r2_traffic_smem.csv python tools/prof_family.py aba_f32 20 > /dev/null 2>&1
RBD_JIT_VARIANT=1 RBD_ONLY=tmem RBD_NO_GATE=1 ncu $M -k regex:rbd_jit_tmem -s 1 --log-file gpurun_out/r2_traffic_tmem.csv python tools/prof_family.py aba_f32 20 > /dev/null 2>&1
RBD_JIT_VARIANT=2 RBD_NO_GATE=1 ncu $M -k regex:rbd_jit_uni -s 1 --log-file gpurun_out/r2_traffic_uni.csv python tools/prof_family.py aba_f32 20 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-other --no-cpu > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E    " | tail -6
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -3 gpurun_out/r2_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null
