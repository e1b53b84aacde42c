#!/bin/bash
# full-set ncu captures of the kernels added late in round 2 (one launch each), condensed to CSV
tag=${1:-r02u}
o=gpurun_out; mkdir -p $o
for pair in "conv3:dense_tile" "sobelgray:sobel_tile" "gaussgray:sep_tile_u8_dp" "gaussrgb:sep_tile_u8_dp" "svd256:jacobi_svd_cluster"; do
  cfg=${pair%%:*}; k=${pair##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -f -o $o/${tag}_ncu_$cfg -k regex:$k -c 1 python tools/gpu_profile_cfg.py $cfg 2 > /dev/null 2>&1
  ncu -i $o/${tag}_ncu_$cfg.ncu-rep --page raw --csv > $o/${tag}_ncu_$cfg.csv 2>/dev/null; rm -f $o/${tag}_ncu_$cfg.ncu-rep
done
ls -la $o | grep ${tag}_ncu
