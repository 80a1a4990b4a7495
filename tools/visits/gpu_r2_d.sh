#!/bin/bash
mkdir -p gpurun_out
date
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
date
timeout 300 python tools/time_small.py 2>&1 | tee gpurun_out/small_latency.txt
date
