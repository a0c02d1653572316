timeout 300 python -m pytest tests/test_gpu_tensorcore.py -x -q -s 2>&1 | tail -40
timeout 300 python scripts/prof_bwd3.py 2>&1 | tail -40
