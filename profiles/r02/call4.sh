#!/bin/bash
set -u
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tree_pipeline.py -q -x > $O/pytest_tree.txt 2>&1; echo "rc=$?" >> $O/pytest_tree.txt
tail -25 $O/pytest_tree.txt
timeout 300 python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_jvp_plain.txt 2>&1; cat $O/c4_jvp_plain.txt | tail -3
