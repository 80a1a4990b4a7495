#!/bin/bash
# One GPU visit: parity tests, bench (+ optional extras passed as arguments).
mkdir -p gpurun_out
date
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
date
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for extra in "$@"; do
  echo "== $extra"; bash -c "$extra" 2>&1 | tail -20
done
date
