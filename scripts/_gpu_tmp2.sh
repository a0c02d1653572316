python -m pytest tests/test_gpu_mapper_pose.py -q -k refinement 2>&1 | tail -3
python - <<'PY'
import sys, torch, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
print(json.dumps(bench.frame_bench(torch.device('cuda', 0)), indent=1))
print(json.dumps({k: v for k, v in bench.render_full_bench(torch.device('cuda', 0)).items() if not k.startswith('_')}, indent=1))
PY
