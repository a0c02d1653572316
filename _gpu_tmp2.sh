timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_scene.py tests/test_gpu_baseline_shapes.py -x -q -s 2>&1 | grep -v "^$" | grep -i "passed\|failed\|error\|max rel err\|loss curve\|weights after\|grid  " | tail -40
timeout 300 python scripts/prof_bwd2.py 2>&1 | tail -32
