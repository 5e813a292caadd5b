#!/bin/bash
set -u
O=gpurun_out/r02q; mkdir -p $O
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread --clock-control none --csv --log-file $O/c3_launches.csv python profiles/run_kernel.py jvp 4096 1 pilz6x2c 100 > $O/ncu_l.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02q/c3_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i0]; ix={k:i for i,k in enumerate(h)}
d=collections.OrderedDict()
for r in rows[i0+1:]:
    if len(r)<len(h): continue
    key=(r[ix['ID']], r[ix['Kernel Name']][:60])
    d.setdefault(key,{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in list(d.items())[-12:]:
    print(k[1], v)
PY
