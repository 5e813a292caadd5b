#!/bin/bash
set -u
O=gpurun_out/r02m; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_tree_chain" -c 1 -o $O/t3 python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_f.log 2>&1
ncu -i $O/t3.ncu-rep --page raw --csv > $O/t3_raw.csv 2>/dev/null
ncu -i $O/t3.ncu-rep --page source --csv > $O/t3_src.csv 2>/dev/null
rm -f $O/t3.ncu-rep; ls -la $O
