python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['phase_ms']); print(d['roofline']['ms_per_launch'], d['roofline']['frac']); print(d['e2e']['value'], d['also']['e2e_autograd_api_rays_per_s'], d['also']['tracking_pose_candidates_per_s'], d['also']['ms_per_frame_640x480'])"
timeout 300 python scripts/prof_e2e.py 2>&1 | head -60
