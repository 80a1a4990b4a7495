#!/bin/bash
# A/B: record rows issued by different warps (PYR_TMA_SPREAD) against one issuing thread
mkdir -p gpurun_out
date
python tools/compare_variants.py save | tail -1
PYR_TOOLS_LIB=libpyrate_b200_spread.so python tools/compare_variants.py check | tail -6
for rep in 1 2; do
for c in "c2_doublegauss 0" "c3_asphere 0" "c1_doublet 1000000"; do
  timeout 300 python tools/time_kernel.py $c 10
  PYR_TOOLS_LIB=libpyrate_b200_spread.so timeout 300 python tools/time_kernel.py $c 10 | sed 's/^/  spread: /'
done; done | tee gpurun_out/timings_spread.txt
date
