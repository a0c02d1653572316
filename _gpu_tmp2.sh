mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracking_query.py tests/test_gpu_multi.py tests/test_manager.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
a=d['also']; print(d['value'], d['ms_per_step'], a['tracking_pose_candidates_per_s'], a['tracking_ms_per_ro_iteration'], a['joint_query_s'], a['ms_per_frame_640x480'] if 'ms_per_frame_640x480' in a else a.get('c4_ms_per_frame_rank0'))
PY
tail -2 gpurun_out/r2_bench_2gpu.err
timeout 300 python scripts/prof_kernels.py ro 5 2>&1 | tail -1 | cut -c1-300
