#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
date
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 2 -c 1 -f -o gpurun_out/prof python tools/profile_target.py c2_doublegauss 0 4 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
date
