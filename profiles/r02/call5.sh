#!/bin/bash
set -u
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tree_pipeline.py -q > $O/pytest_tree.txt 2>&1; echo "rc=$?" >> $O/pytest_tree.txt
tail -8 $O/pytest_tree.txt
python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_jvp_plain.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_launches.csv python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_tree" -s 3 -c 3 -o $O/r02_c4_tree python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/ncu_f.log 2>&1
ncu -i $O/r02_c4_tree.ncu-rep --page raw --csv > $O/r02_c4_tree_raw.csv 2>/dev/null
ncu -i $O/r02_c4_tree.ncu-rep --page source --csv > $O/r02_c4_tree_src.csv 2>/dev/null
ls -la $O; rm -f $O/r02_c4_tree.ncu-rep; grep k_tree $O/c4_launches.csv | cut -c1-200 | tail -6
