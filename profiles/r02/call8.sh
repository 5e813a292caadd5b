#!/bin/bash
set -u
O=gpurun_out/r02h; mkdir -p $O
timeout 2000 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -12 $O/pytest_gpu.txt
