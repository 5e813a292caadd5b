#!/bin/bash
set -u
O=gpurun_out/r02t; mkdir -p $O
for v in 2_2 1_2 2_1 1_1; do
  if [ $v = default ]; then unset MPCF_LIB; else export MPCF_LIB=$PWD/gpurun_in/libmpcf_$v.so; fi
  echo "variant $v"; python profiles/run_kernel.py jvp 1024 3 humanoid37 40 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l_$v.csv -k regex:"k_tree_stages|k_tree_derivs" -c 2 python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > /dev/null 2>&1
  grep k_tree $O/l_$v.csv | awk -F'","' '{print substr($5,1,24), $NF}'
done
