#!/bin/bash
# round-2 GPU validation (2 GPUs): multi-GPU parity through zb_shard_*, the GPU suite, bench at N = 1 and 2, launch lists, all configs
tag=${1:-r02c}
o=gpurun_out; mkdir -p $o
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/gpu_shard_check.py > $o/${tag}_shard_check_n2.log 2>&1; echo "shard check rc=$?"; grep -E "MISMATCH|shard check|Error|error" $o/${tag}_shard_check_n2.log | head
timeout 300 $TR --nproc-per-node 2 --master-port 29513 tools/gpu_shard_diag.py 8192 > $o/${tag}_shard_diag.log 2>&1; grep "rank" $o/${tag}_shard_diag.log
timeout 300 $TR --nproc-per-node 2 --master-port 29514 tools/gpu_shard_diag.py 4096 >> $o/${tag}_shard_diag.log 2>&1; grep "rank" $o/${tag}_shard_diag.log | tail -2
timeout 1500 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $o/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 100 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 100 > $o/${tag}_bench_n2.json 2> $o/${tag}_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $o/${tag}_bench_n2.json; tail -3 $o/${tag}_bench_n2.err
for cfg in fdm rotate gauss8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_launches_$cfg.csv python tools/gpu_profile_cfg.py $cfg 4 > /dev/null 2>&1
  echo "== $cfg"; grep -h "gpu__time_duration" $o/${tag}_launches_$cfg.csv | awk -F'","' '{print $5, $(NF)}' | tail -8
done
timeout 600 python tools/gpu_bench_all.py > $o/${tag}_bench_all.log 2>&1; tail -40 $o/${tag}_bench_all.log
