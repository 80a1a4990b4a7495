#!/bin/bash
# parity + timings: real-valued fast path of the crystal kernel
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for c in "c4_anisotropic 1000000" "x4_biaxial 1000000" "x8_crystal_mirror 1000000" "x5_degenerate 1000000"; do timeout 300 python tools/time_kernel.py $c 10 2>&1 | tail -1; done | tee gpurun_out/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_complex -s 1 -c 1 -f -o gpurun_out/prof_r02o_c4 python tools/profile_target.py c4_anisotropic 1000000 4 mem > gpurun_out/ncu_c4.log 2>&1; tail -1 gpurun_out/ncu_c4.log
date
