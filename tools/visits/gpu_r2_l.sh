#!/bin/bash
# parity + timings: approximate Newton seed (asphere), branch-free sqrt / division in the crystal kernel; variants
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for c in "c2_doublegauss 0" "c3_asphere 0" "c4_anisotropic 1000000" "x4_biaxial 1000000" "x2_xypoly 4000000"; do timeout 300 python tools/time_kernel.py $c 10; done | tee gpurun_out/timings.txt
echo "variant: asphere kernel 128 threads x 4 CTAs/SM" | tee -a gpurun_out/timings.txt
PYR_TOOLS_LIB=libpyrate_b200_a128.so timeout 300 python tools/time_kernel.py c3_asphere 0 10 | tee -a gpurun_out/timings.txt
echo "variant: crystal kernel 3 CTAs/SM (168 registers, spills)" | tee -a gpurun_out/timings.txt
PYR_TOOLS_LIB=libpyrate_b200_c4m3.so timeout 300 python tools/time_kernel.py c4_anisotropic 1000000 10 | tee -a gpurun_out/timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_complex -s 1 -c 1 -f -o gpurun_out/prof_r02l_c4 python tools/profile_target.py c4_anisotropic 1000000 4 mem > gpurun_out/ncu_c4.log 2>&1; tail -1 gpurun_out/ncu_c4.log
date
