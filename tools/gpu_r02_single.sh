#!/bin/bash
# round-2 single-GPU validation: the GPU suite, bench N = 1, launch lists of fdm / rotate / gauss8, all-config timings
tag=${1:-r02d}
o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $o/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 100 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
for cfg in fdm rotate gauss8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_launches_$cfg.csv python tools/gpu_profile_cfg.py $cfg 4 > /dev/null 2>&1
  echo "== $cfg"; grep -h "gpu__time_duration" $o/${tag}_launches_$cfg.csv | awk -F'","' '{print $5, $(NF)}' | tail -8
done
timeout 600 python tools/gpu_bench_all.py > $o/${tag}_bench_all.log 2>&1; tail -40 $o/${tag}_bench_all.log
