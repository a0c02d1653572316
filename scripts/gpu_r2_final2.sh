#!/bin/bash
# Round-2 closing evidence batch: GPU tests, smoke, both bench arms, launch list, ncu --set full of the marching-cubes kernels,
# compute-sanitizer memcheck of the new kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gpu_pytest.log
tail -4 gpurun_out/r2_gpu_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2>> gpurun_out/r2_bench.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 150 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'mc_nodes_kernel|mc_classify_kernel|mc_emit_kernel|mc_resolve_kernel' -s 4 -c 4 -o gpurun_out/r2_mc python scripts/prof_mcubes.py 512 > gpurun_out/r2_ncu_mc.log 2>&1; echo "ncu mc rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_marching_cubes.py tests/test_mesher.py -q -m gpu -k "golden or edge or chains or visibility" > gpurun_out/r2_sanitizer_mcubes.log 2>&1; echo "sanitizer rc=$?"
tail -4 gpurun_out/r2_sanitizer_mcubes.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], r['value'], d['roofline']['frac'], d['clocks'])
PY
ls -la gpurun_out | tail -8
