#!/bin/bash
set -u
O=gpurun_out/r02final; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_tree" -s 10 -c 5 -o $O/r02_c4_tc python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/c4_ncu.log 2>&1
MPCF_TREE_CHAIN=scalar ncu --set full --clock-control none -k regex:"k_tree" -s 8 -c 4 -o $O/r02_c4_tree python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/c4s_ncu.log 2>&1
for r in r02_c4_tc r02_c4_tree; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep; ls -la $O | grep c4
