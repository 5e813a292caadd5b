#!/bin/bash
# Round-2 evidence run (one GPU): full -m gpu suite, smoke, default bench (C2 headline + C1/C3/C4/C5), reference arm, the ncu
# launch list of the bench command, ncu --set full of the pipeline kernels of C2 / C3 / C4, raw timings of the other entries.
set -u
O=gpurun_out/r02final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --only-headline > $O/bench_ncu.log 2>&1
python profiles/run_configs.py > $O/run_configs.txt 2>&1
python profiles/time_frames.py > $O/time_frames.txt 2>&1
python profiles/run_kernel.py jvp 16384 2 > $O/c2_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_step_stages|k_stage_derivs|k_chain_rule" -s 6 -c 3 -o $O/r02_c2_jvp python profiles/run_kernel.py jvp 16384 2 > $O/c2_ncu.log 2>&1
python profiles/run_kernel.py jvp 4096 1 pilz6x2c 100 > $O/c3_plain.txt 2>&1 &&
ncu --set full --clock-control none -k regex:"k_step_stages|k_stage_derivs|k_chain_rule|k_couple|k_forest_fill" -s 9 -c 9 -o $O/r02_c3_jvp python profiles/run_kernel.py jvp 4096 1 pilz6x2c 100 > $O/c3_ncu.log 2>&1
python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tree" -s 10 -c 5 -o $O/r02_c4_tc python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/c4_ncu.log 2>&1
# the DFMA variant of the chain kernel (no padding): its executed-FLOP count is the work model of C4
MPCF_TREE_CHAIN=scalar python profiles/run_kernel.py jvp 1024 2 humanoid37 40 > $O/c4_scalar_plain.txt 2>&1 &&
MPCF_TREE_CHAIN=scalar ncu --set full --clock-control none -k regex:"k_tree" -s 8 -c 4 -o $O/r02_c4_tree python profiles/run_kernel.py jvp 1024 1 humanoid37 40 > $O/c4s_ncu.log 2>&1
for r in r02_c2_jvp r02_c3_jvp r02_c4_tc r02_c4_tree; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
rm -f $O/*.ncu-rep
ls -la $O; cat $O/c2_plain.txt $O/c3_plain.txt $O/c4_plain.txt $O/c4_scalar_plain.txt
