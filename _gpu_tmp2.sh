timeout 900 python -m pytest tests/test_gpu_scene.py tests/test_gpu_tensorcore.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_bwd2.py 2>&1 | head -7
