mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2b_bench_8gpu.json 2> gpurun_out/r2b_bench_8gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_8gpu.json').read().strip().splitlines()[-1])
a=d['also']; print(d['value'], d['ms_per_step'], d['e2e']['value'], a['tracking_pose_candidates_per_s'], a['joint_query_s'], a['joint_query_submap_sharded_s'], a.get('c4_ms_per_frame_rank0'))
PY
