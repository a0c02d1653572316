#!/bin/bash
# Round-2 evidence batch: GPU tests, the bench line, launch list and ncu --set full captures of every dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -6 gpurun_out/r2_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err; echo "ref rc=$?"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 150 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc3_kernel|field_bwd_tc2_kernel|adam_pair_kernel' -s 9 -c 3 -o gpurun_out/r2_map python scripts/prof_kernels.py map 4 > gpurun_out/r2_ncu_map.log 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc3_kernel' -s 2 -c 1 -o gpurun_out/r2_ro python scripts/prof_kernels.py ro 1 > gpurun_out/r2_ncu_ro.log 2>&1
$NCU --set full --import-source on -k regex:'field_fwd_tc3_kernel' -s 5 -c 1 -o gpurun_out/r2_jq python scripts/prof_kernels.py jq > gpurun_out/r2_ncu_jq.log 2>&1
$NCU --set full --import-source on -k regex:'field_bwd_tc_kernel' -s 2 -c 1 -o gpurun_out/r2_go python scripts/prof_kernels.py go 3 > gpurun_out/r2_ncu_go.log 2>&1
timeout 300 python scripts/prof_bwd2.py > gpurun_out/r2_prof_bwd2.txt 2>&1
timeout 300 python scripts/prof_cta.py > gpurun_out/r2_prof_cta.txt 2>&1
timeout 300 python scripts/prof_dropin.py > gpurun_out/r2_prof_dropin.txt 2>&1
timeout 300 python scripts/ab_dynamic_tiles.py > gpurun_out/r2_ab_dynamic_tiles.txt 2>&1
ls -la gpurun_out/ | tail -20
