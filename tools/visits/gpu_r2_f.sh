#!/bin/bash
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
date
for c in "c5_grin 1000000" "c5_grin 12500000" "c4_anisotropic 1000000" "x4_biaxial 1000000" "x8_crystal_mirror 1000000"; do timeout 300 python tools/time_kernel.py $c 10; done 2>&1 | tee gpurun_out/timings.txt
echo "-- tools build (PYR_C4_MINB=2)"; PYR_LEAN_VARIANT=1 timeout 300 python tools/time_kernel.py c4_anisotropic 1000000 10 2>&1 | tee -a gpurun_out/timings.txt
for cfg in "c5_grin:1000000:trace_real:1:c5" "c4_anisotropic:1000000:trace_complex:1:c4"; do
  cfg=${cfg//:/ }
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/prof_r02f_$5 python tools/profile_target.py $1 $2 4 > gpurun_out/ncu_$5.log 2>&1; tail -1 gpurun_out/ncu_$5.log
done
date
