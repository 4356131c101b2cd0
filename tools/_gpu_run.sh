set -x
python tools/jit_check.py atlas dynamics f32 20
python tools/jit_check.py atlas id f32 20
python tools/jit_check.py iiwa14 dynamics f32 20
python tools/jit_check.py valkyrie dynamics f32 20
python tools/jit_sweep.py 24 31
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1
