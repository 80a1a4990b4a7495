#!/bin/bash
for v in 0 50; do PYR_LEAN_VARIANT=$v timeout 300 python tools/time_kernel.py c2_doublegauss 0 20; done
timeout 300 python tools/time_kernel.py c2_doublegauss 0 10 1
timeout 300 python tools/time_kernel.py c1_doublet 1000000 10
timeout 300 python tools/time_kernel.py x1_tilted 4000000 10
timeout 300 python tools/time_kernel.py c3_asphere 0 10
timeout 300 python tools/time_kernel.py x2_xypoly 4000000 10
timeout 300 python tools/time_kernel.py c5_grin 1000000 5
timeout 300 python tools/time_kernel.py c4_anisotropic 1000000 5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 2 -c 1 -f -o gpurun_out/prof_v5 python tools/profile_target.py c2_doublegauss 0 4 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
