#!/bin/bash
# rotate tile kernel: parity tests, timing against the gather kernel at three angles, launch list + one full ncu capture
tag=${1:-r02f}
o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_warp.py -m gpu -q -x > $o/${tag}_pytest_warp.log 2>&1; echo "pytest rc=$?"; tail -15 $o/${tag}_pytest_warp.log
for a in 0.7853981633974483 0.1 1.3; do timeout 120 python tools/gpu_rotate_check.py 128 $a >> $o/${tag}_rotate_check.log 2>&1; done; cat $o/${tag}_rotate_check.log
timeout 300 ncu --set full --clock-control none --import-source on -f -o $o/${tag}_ncu_rotate_tile -k regex:rotate_tile -c 1 python tools/gpu_profile_cfg.py rotate 2 > /dev/null 2>&1
ncu -i $o/${tag}_ncu_rotate_tile.ncu-rep --page raw --csv > $o/${tag}_ncu_rotate_tile.csv 2>/dev/null
ncu -i $o/${tag}_ncu_rotate_tile.ncu-rep --page source --csv > $o/${tag}_ncu_rotate_tile_source.csv 2>/dev/null; rm -f $o/${tag}_ncu_rotate_tile.ncu-rep
ls -la $o | tail -5
