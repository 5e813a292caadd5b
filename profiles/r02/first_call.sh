#!/bin/bash
# Round-2 first GPU call: evidence that round 1 left out (raw output of run_configs / time_frames, ncu of the
# run-time-tree kernels, FLOP counters of the C3 pipeline, K3 CPW A/B).  Everything lands in gpurun_out/r02/.
set -u
O=gpurun_out/r02; mkdir -p $O
python profiles/run_configs.py > $O/run_configs.txt 2>&1
python profiles/time_frames.py > $O/time_frames.txt 2>&1
python profiles/run_kernel.py jvp 16384 2 > $O/k3_cpw1.txt 2>&1
MPCF_K3_CPW=2 python profiles/run_kernel.py jvp 16384 2 > $O/k3_cpw2.txt 2>&1
# C4 (humanoid37, generic64): plain runs first, then ncu
python profiles/run_kernel.py step 8192 2 humanoid37 40 > $O/c4_step_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:generic_kernel -s 1 -c 1 -o $O/r02_c4_step \
    python profiles/run_kernel.py step 8192 1 humanoid37 40 > $O/c4_step_ncu.log 2>&1
python profiles/run_kernel.py jvp 64 1 humanoid37 40 > $O/c4_jvp_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:generic_kernel -s 1 -c 1 -o $O/r02_c4_jvp \
    python profiles/run_kernel.py jvp 64 1 humanoid37 40 > $O/c4_jvp_ncu.log 2>&1
python profiles/run_kernel.py rnea 8192 2 humanoid37 40 > $O/c4_rnea_plain.txt 2>&1 &&
ncu --set full --clock-control none -k regex:generic_kernel -s 1 -c 1 -o $O/r02_c4_rnea \
    python profiles/run_kernel.py rnea 8192 1 humanoid37 40 > $O/c4_rnea_ncu.log 2>&1
# C3 (forest12x6): FLOP counters + traffic of its pipeline kernels
python profiles/run_kernel.py jvp 4096 1 pilz6x2 100 > $O/c3_jvp_plain.txt 2>&1 &&
ncu --set full --clock-control none -k regex:"k_step_stages|k_stage_derivs|k_chain_rule" -s 6 -c 6 -o $O/r02_c3_jvp \
    python profiles/run_kernel.py jvp 4096 1 pilz6x2 100 > $O/c3_jvp_ncu.log 2>&1
# keep the merge-back under 64 MiB: export the raw / source pages here and drop the big reports
for r in r02_c4_step r02_c4_jvp r02_c4_rnea r02_c3_jvp; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
[ -f $O/r02_c4_step.ncu-rep ] && ncu -i $O/r02_c4_step.ncu-rep --page source --csv > $O/r02_c4_step_src.csv 2>/dev/null
rm -f $O/r02_c3_jvp.ncu-rep $O/r02_c4_jvp.ncu-rep $O/r02_c4_rnea.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
ls -la $O
