#!/bin/bash
set -u
python profiles/run_kernel.py jvp 16384 3 pilz6 100 2>&1 | tail -1
MPCF_K3_CPW=2 python profiles/run_kernel.py jvp 16384 3 pilz6 100 2>&1 | tail -1
