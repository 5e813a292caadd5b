#!/bin/bash
set -u
O=gpurun_out/r02r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
tail -5 $O/pytest.txt
python profiles/run_kernel.py jvp 16384 3 pilz6 100 2>&1 | tail -1
python profiles/run_kernel.py jvp 4096 3 pilz6x2c 100 2>&1 | tail -1
python bench.py --no-cpu --only-headline --steps 5 --warmup 3 > $O/bench_headline.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r/bench_headline.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'].get('frac'), d['roofline'].get('per_kernel_ms', d['roofline'].get('kernels')))
PY
