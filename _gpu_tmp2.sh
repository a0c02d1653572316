timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tracking_query.py -x -q 2>&1 | grep -v "^$" | tail -25
