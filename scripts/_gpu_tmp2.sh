python -m pytest tests/test_gpu_tensorcore.py -q -s -k "role_split" 2>&1 | grep -E "passed|failed|Error|error|grid " | head -20
timeout 300 python scripts/prof_bwd2.py 2>&1 | tail -32
