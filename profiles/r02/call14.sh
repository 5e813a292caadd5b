#!/bin/bash
set -u
O=gpurun_out/r02n; mkdir -p $O
for d in 0 1 2 4 3 7; do
echo "dbg=$d" >> $O/dbg.txt
MPCF_TC_DBG=$d python profiles/run_kernel.py jvp 1024 3 humanoid37 40 >> $O/dbg.txt 2>&1
done
cat $O/dbg.txt
