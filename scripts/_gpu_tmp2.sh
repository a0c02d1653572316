set -x
python -m pytest tests/test_gpu_tensorcore.py -q -s -k "role_split" 2>&1 | grep -E "vs oracle|passed|failed|Error|error" | head -60
timeout 300 python scripts/prof_bwd2.py 2>&1 | tail -30
python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_tracking_query.py -x -q -s 2>&1 | tail -30
timeout 300 python scripts/run_ro.py 5 2>&1 | tail -3
