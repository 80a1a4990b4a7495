#!/bin/bash
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
date
for c in "c5_grin 1000000" "c5_grin 12500000" "c3_asphere 0" "c4_anisotropic 1000000"; do timeout 300 python tools/time_kernel.py $c 10; done 2>&1 | tee gpurun_out/timings.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 1 -c 1 -f -o gpurun_out/prof_r02e_c5 python tools/profile_target.py c5_grin 1000000 4 > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
date
