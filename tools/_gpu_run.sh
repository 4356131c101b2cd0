set -x
python tools/jit_check.py atlas dynamics f32 20
RBD_ONLY=smem python tools/jit_check.py atlas dynamics f32 20
RBD_SMEM_BLOCKS=4 python tools/jit_check.py atlas dynamics f32 20
python tools/jit_check.py atlas id f32 20
python tools/jit_check.py iiwa14 dynamics f32 20
python tools/jit_check.py iiwa14 id f32 20
python tools/jit_check.py valkyrie dynamics f32 20
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
