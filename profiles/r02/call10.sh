#!/bin/bash
set -u
O=gpurun_out/r02j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tree_pipeline.py -q -x > $O/pytest_tree.txt 2>&1; echo "rc=$?" >> $O/pytest_tree.txt
tail -15 $O/pytest_tree.txt
python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_jvp_plain.txt 2>&1; cat $O/c4_jvp_plain.txt
MPCF_TREE_CHAIN=scalar python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_jvp_scalar.txt 2>&1; cat $O/c4_jvp_scalar.txt
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64_op_dmma.sum --clock-control none --csv --log-file $O/c4_launches.csv -k regex:k_tree python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_l.log 2>&1
grep k_tree $O/c4_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-140 | tail -12
