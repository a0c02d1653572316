timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -s 2>&1 | grep -v "^$" | grep -i "passed\|failed\|error\|grid  \|assert" | tail -12
timeout 300 python scripts/prof_bwd2.py 2>&1 | head -30
