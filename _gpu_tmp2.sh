timeout 900 python -m pytest tests/test_gpu_scene.py tests/test_gpu_tracking_query.py tests/test_gpu_tensorcore.py tests/test_gpu_mapper_pose.py -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_cta.py 2>&1 | tail -4
timeout 300 python scripts/ab_dynamic_tiles.py 2>&1 | grep "dynamic tiles 1"
timeout 300 python - <<'PY' 2>&1 | tail -3
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
r = bench.frame_bench(torch.device("cuda", 0), frames=6)
print({k: v for k, v in r.items() if k != "frame_shape"})
PY
