#!/bin/bash
# parity + timings: GRIN loop (512-entry exp table / quartic, cylinder-boundary copy, one z drift per step)
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for c in "c5_grin 1000000" "c5_grin 12500000" "x17_sech_rod 1000000"; do timeout 300 python tools/time_kernel.py $c 10; done 2>&1 | tail -4 | tee gpurun_out/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 1 -c 1 -f -o gpurun_out/prof_r02m_c5 python tools/profile_target.py c5_grin 1000000 4 mem > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
date
