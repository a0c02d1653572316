timeout 300 python scripts/prof_kernels.py jq512 2>&1 | tail -1
timeout 300 python scripts/prof_kernels.py jq512 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_tracking_query.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -3
