python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_encodings.py tests/test_gpu_baseline_shapes.py -q 2>&1 | tail -3
timeout 300 python scripts/prof_bwd2.py 2>&1 | grep -E "bwd impl|chain"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['roofline']['ms_per_launch']); a=d['also']; print(a['tracking_ms_per_ro_iteration'], a['joint_query_s'], a['render_full_img_ms'], a['ms_per_frame_640x480'], a['pose_refinement_10_iterations_ms'])"
