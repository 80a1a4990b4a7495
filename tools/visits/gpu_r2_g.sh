#!/bin/bash
mkdir -p gpurun_out
date
timeout 900 python -m pytest tests -m gpu -q -k "grin or merit or fused" 2>&1 | tail -4
echo "-- default build"; timeout 300 python tools/time_kernel.py c5_grin 12500000 6 2>&1 | tee gpurun_out/timings_grin.txt
for v in r1_m2 r1_m3 r1_m4 r2_m1 r2_m3; do
  echo "-- variant $v"; PYR_TOOLS_LIB=libpyr_grin_$v.so timeout 300 python tools/time_kernel.py c5_grin 12500000 6 2>&1 | tee -a gpurun_out/timings_grin.txt
done
date
