#!/bin/bash
# round 2, visit A: parity suite + first timings of the generated-bundle path
mkdir -p gpurun_out
date
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
date
for c in "c2_doublegauss 0" "c3_asphere 0" "c1_doublet 1000000" "c5_grin 1000000"; do timeout 300 python tools/time_gen.py $c 10 2>&1 | tail -6; done | tee gpurun_out/time_gen.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
date
