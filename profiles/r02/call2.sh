#!/bin/bash
# GPU call 2: full -m gpu test suite, smoke, default bench (all configs), reference arm
set -u
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
tail -3 $O/pytest_gpu.txt; tail -2 $O/smoke.txt; tail -3 $O/bench_default.err; head -c 600 $O/bench_default.json
