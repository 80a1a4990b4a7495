#!/bin/bash
# round 2, visit B: parity suite on the single-launch crystal path, timings, ncu captures
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
date
for c in "c2_doublegauss 0" "c3_asphere 0" "c4_anisotropic 1000000" "c5_grin 1000000" "x4_biaxial 1000000" "x16_cylinder 4000000"; do timeout 300 python tools/time_kernel.py $c 10; done 2>&1 | tee gpurun_out/timings.txt
CAPTURES=${CAPTURES:-"c3_asphere:0:trace_real:1:c3 c4_anisotropic:1000000:trace_complex:1:c4 c5_grin:1000000:trace_real:1:c5"}
for cfg in $CAPTURES; do
  cfg=${cfg//:/ }
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/prof_r02b_$5 python tools/profile_target.py $1 $2 4 > gpurun_out/ncu_$5.log 2>&1; tail -1 gpurun_out/ncu_$5.log
done
date
