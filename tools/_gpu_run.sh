set -x
RBD_ONLY=smem ncu --set full --clock-control none --import-source on -k regex:rbd_jit -s 1 -c 1 -o gpurun_out/r2_jit_packed_smem -f python tools/prof_one.py 262144 2>&1 | tail -3
RBD_JIT_CONVOY=0 RBD_JIT_PACKED=0 RBD_ONLY=smem RBD_JIT_CACHE=/tmp/jc ncu --set full --clock-control none --import-source on -k regex:rbd_jit -s 1 -c 1 -o gpurun_out/r2_jit_scalar_noconvoy_smem -f python tools/prof_one.py 262144 2>&1 | tail -3
