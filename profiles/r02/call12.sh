#!/bin/bash
set -u
O=gpurun_out/r02l; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_tree_chain" -c 2 -o $O/t3 python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_f.log 2>&1
ncu -i $O/t3.ncu-rep --page raw --csv > $O/t3_raw.csv 2>/dev/null
rm -f $O/t3.ncu-rep; ls -la $O
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
