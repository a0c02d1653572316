timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_tensorcore.py tests/test_gpu_baseline_shapes.py tests/test_gpu_mapper_pose.py -x -q 2>&1 | tail -15
timeout 300 python scripts/prof_dropin.py 2>&1 | tail -12
