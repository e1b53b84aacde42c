#!/bin/bash
# Collects the single-GPU evidence profiles/ is built from: GPU parity log, bench lines (both arms), all-config timings, the ncu
# launch list of the bench command and full-set captures of the hot kernels.  usage: tools/gpu_round_artifacts.sh <tag>
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $o/${tag}_pytest_gpu.log
timeout 600 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"; cut -c1-300 $o/${tag}_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err; cut -c1-300 $o/${tag}_bench_reference.json
timeout 600 python tools/gpu_bench_all.py > $o/${tag}_bench_all.log 2>&1; cp $o/all_configs.json $o/${tag}_all_configs.json 2>/dev/null; tail -30 $o/${tag}_bench_all.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -o $o/${tag}_ncu_fused -k regex:fused_sep -c 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i $o/${tag}_ncu_fused.ncu-rep --page raw --csv > $o/${tag}_ncu_fused.csv 2>/dev/null; rm -f $o/${tag}_ncu_fused.ncu-rep
for cfg in fdm rotate gauss8 bicubic gemm boxblur; do
  timeout 300 ncu --set full --clock-control none --import-source on -f -o $o/${tag}_ncu_$cfg -k regex:'rotate|box_|xtx_tf32|resize|fdm|moments|fused_sep_rgba8' -c 4 python tools/gpu_profile_cfg.py $cfg 1 > /dev/null 2>&1
  ncu -i $o/${tag}_ncu_$cfg.ncu-rep --page raw --csv > $o/${tag}_ncu_$cfg.csv 2>/dev/null; rm -f $o/${tag}_ncu_$cfg.ncu-rep   # (only the condensed CSV travels back: 64 MiB limit)
done
timeout 200 python tools/gpu_gemm_tc_check.py > $o/${tag}_gemm_tc_check.log 2>&1; tail -6 $o/${tag}_gemm_tc_check.log
ls -la $o | tail -20
