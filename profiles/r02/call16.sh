#!/bin/bash
set -u
O=gpurun_out/r02p; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -k "forest or dual or coupled or pilz6x2 or parity or fullsize" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -5 $O/pytest.txt
python profiles/run_kernel.py jvp 4096 3 pilz6x2c 100 2>&1 | tail -2
python profiles/run_kernel.py jvp 4096 3 pilz6x2 100 2>&1 | tail -2
