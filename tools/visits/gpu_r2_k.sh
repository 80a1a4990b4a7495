#!/bin/bash
# parity + timings after the joint-Newton / sqrt-free asphere and the leaner GRIN stage
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for c in "c2_doublegauss 0" "c3_asphere 0" "c5_grin 1000000" "x2_xypoly 4000000" "x6_biconic 4000000" "x1_tilted 4000000" "c4_anisotropic 1000000"; do timeout 300 python tools/time_kernel.py $c 10; done | tee gpurun_out/timings.txt
for c in "c3_asphere 0"; do timeout 300 python tools/time_gen.py $c 10 2>&1 | head -2; done | tee gpurun_out/time_gen.txt
for cfg in "c3_asphere:0:trace_real:1:c3:mem" "c5_grin:1000000:trace_real:1:c5:mem"; do
  cfg=${cfg//:/ }
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/prof_r02k_$5 python tools/profile_target.py $1 $2 4 $6 > gpurun_out/ncu_$5.log 2>&1; tail -1 gpurun_out/ncu_$5.log
done
date
