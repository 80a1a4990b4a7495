#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "asphere or fixture or random or larger" 2>&1 | tail -3
for rep in 1 2; do timeout 300 python tools/time_kernel.py c3_asphere 0 10; done | tee gpurun_out/timings_c3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 1 -c 1 -f -o gpurun_out/prof_r02p_c3 python tools/profile_target.py c3_asphere 0 4 mem > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
