#!/bin/bash
# round-2 first GPU call (2 GPUs): multi-GPU parity through zb_shard_*, the GPU suite, bench at N = 1 and 2
o=gpurun_out; mkdir -p $o
nvidia-smi -L > $o/r02b_gpus.log 2>&1; nvidia-smi topo -m >> $o/r02b_gpus.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/gpu_shard_check.py > $o/r02b_shard_check_n2.log 2>&1; echo "shard check rc=$?"; tail -40 $o/r02b_shard_check_n2.log
timeout 300 python tools/gpu_shard_check.py --quick > $o/r02b_shard_check_n1.log 2>&1; echo "shard check n1 rc=$?"; tail -5 $o/r02b_shard_check_n1.log
timeout 1500 python -m pytest tests -m gpu -x -q > $o/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $o/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 100 > $o/r02b_bench_n1.json 2> $o/r02b_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-1500 $o/r02b_bench_n1.json; tail -5 $o/r02b_bench_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 100 > $o/r02b_bench_n2.json 2> $o/r02b_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-2500 $o/r02b_bench_n2.json; tail -5 $o/r02b_bench_n2.err
