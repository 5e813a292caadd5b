#!/bin/bash
set -u
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tree_pipeline.py -q -x > $O/pytest_tree.txt 2>&1; echo "rc=$?" >> $O/pytest_tree.txt
tail -4 $O/pytest_tree.txt
for d in 0 1 2 3; do
echo "dbg=$d" >> $O/dbg.txt
MPCF_TC_DBG=$d python profiles/run_kernel.py jvp 1024 3 humanoid37 40 >> $O/dbg.txt 2>&1
done
cat $O/dbg.txt
MPCF_TREE_CHAIN=scalar python profiles/run_kernel.py jvp 1024 2 humanoid37 40
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_launches.csv -k regex:k_tree python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_l.log 2>&1
grep k_tree $O/c4_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | head -4
