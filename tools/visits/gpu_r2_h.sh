#!/bin/bash
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
for c in "c4_anisotropic 1000000" "c5_grin 1000000"; do timeout 300 python tools/time_kernel.py $c 10; done 2>&1 | tee gpurun_out/timings.txt
date
