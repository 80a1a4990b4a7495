#!/bin/bash
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
date
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
date
