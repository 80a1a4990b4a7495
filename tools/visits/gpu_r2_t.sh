#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "grin or c5 or fixture or plugin or history or lockstep" 2>&1 | tail -3
for c in "c5_grin 1000000" "c5_grin 12500000"; do timeout 300 python tools/time_kernel.py $c 10; done | tee gpurun_out/timings_t.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 1 -c 1 -f -o gpurun_out/prof_r02t_c5 python tools/profile_target.py c5_grin 1000000 4 mem > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
