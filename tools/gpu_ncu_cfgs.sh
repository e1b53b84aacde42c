#!/bin/bash
# usage: tools/gpu_ncu_cfgs.sh cfg1 cfg2 ...   -> gpurun_out/ncu_<cfg>.ncu-rep + .csv (raw page), one full-set capture per kernel launch of the 2nd repetition
mkdir -p gpurun_out
for cfg in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/ncu_$cfg \
     -k regex:'rotate|sat_|box_|resize|xtx|fdm|moments|conv|warp|canny|order_|ssim|diff_sums|motion_' -c 8 python tools/gpu_profile_cfg.py $cfg 1 > gpurun_out/ncu_$cfg.log 2>&1
  ncu -i gpurun_out/ncu_$cfg.ncu-rep --page raw --csv > gpurun_out/ncu_$cfg.csv 2>/dev/null
  tail -2 gpurun_out/ncu_$cfg.log
done
