timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
PY
