python tools/jit_check.py atlas dynamics f32 20
python tools/jit_check.py atlas id f32 20
python tools/jit_check.py iiwa14 dynamics f32 20
python tools/jit_check.py valkyrie dynamics f32 20
python tools/jit_sweep.py 10 18 31
