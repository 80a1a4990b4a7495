#!/bin/bash
# A/B: plain path with static tiles vs dynamic in-order tile hand-out (PYR_DYNAMIC_TILES)
mkdir -p gpurun_out
python tools/compare_variants.py save | tail -1
PYR_TOOLS_LIB=libpyrate_b200_dyn.so python tools/compare_variants.py check | tail -6
for rep in 1 2; do
for c in "c2_doublegauss 0" "c3_asphere 0"; do
  timeout 300 python tools/time_kernel.py $c 10
  PYR_TOOLS_LIB=libpyrate_b200_dyn.so timeout 300 python tools/time_kernel.py $c 10 | sed 's/^/  dyn: /'
done; done | tee gpurun_out/timings_dyn.txt
PYR_TOOLS_LIB=libpyrate_b200_dyn.so timeout 300 python tools/time_gen.py c2_doublegauss 0 10 2>&1 | head -1 | sed 's/^/  dyn: /'
timeout 300 python tools/time_gen.py c2_doublegauss 0 10 2>&1 | head -1
