#!/bin/bash
# multi-GPU validation: parity through zb_shard_* (tools/gpu_shard_check.py), then bench.py at N ranks (weak + strong + C4 / C5 sharded, in-run parity)
n=${1:-2}; tag=${2:-r02m}
o=gpurun_out; mkdir -p $o
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $n --master-port 29511 tools/gpu_shard_check.py > $o/${tag}_shard_check_n$n.log 2>&1; echo "shard check rc=$?"; grep -E "MISMATCH|shard check|Error|error|ok" $o/${tag}_shard_check_n$n.log | head -40
timeout 600 $TR --nproc-per-node $n --master-port 29512 bench.py --gpus $n --steps 100 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "bench n$n rc=$?"; cat $o/${tag}_bench_n$n.json; tail -3 $o/${tag}_bench_n$n.err
