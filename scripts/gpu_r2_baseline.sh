#!/bin/bash
# Round-2 first GPU batch: GPU tests, a bench line, launch list and ncu --set full captures of the kernels that
# round 1 left without a record (forward, RandomOptimizer field query, joint-query field query, pose-gradient backward).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.csv
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc3_kernel|field_bwd_tc_kernel|adam_pair_kernel' -s 9 -c 3 -o gpurun_out/r2a_map python scripts/prof_kernels.py map 4 > gpurun_out/r2a_ncu_map.log 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc3_kernel' -s 2 -c 1 -o gpurun_out/r2a_ro python scripts/prof_kernels.py ro 1 > gpurun_out/r2a_ncu_ro.log 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc2_kernel' -s 20 -c 1 -o gpurun_out/r2a_jq python scripts/prof_kernels.py jq > gpurun_out/r2a_ncu_jq.log 2>&1
$NCU --set full --import-source on -k regex:'field_bwd_kernel' -s 1 -c 1 -o gpurun_out/r2a_go python scripts/prof_kernels.py go 3 > gpurun_out/r2a_ncu_go.log 2>&1
ls -la gpurun_out/
