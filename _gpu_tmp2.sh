timeout 600 python -m pytest tests/test_gpu_mapper_pose.py -x -q -s 2>&1 | grep -v "^$" | grep -i "passed\|failed\|error\|pose refinement\|assert" | tail -12
timeout 300 python - <<'PY' 2>&1 | tail -6
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
r = bench.frame_bench(torch.device("cuda", 0), frames=6)
print({k: v for k, v in r.items() if k != "frame_shape"})
PY
