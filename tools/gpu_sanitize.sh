#!/bin/bash
# compute-sanitizer passes over every kernel family on small bundles
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for cfg in "c2_doublegauss 20000" "c3_asphere 9000" "c5_grin 3000" "c4_anisotropic 3000" "x1_tilted 7001"; do
    set -- $cfg
    echo "== $tool $1 $2"
    timeout 280 compute-sanitizer --tool $tool --print-limit 5 python tools/profile_target.py $1 $2 1 2>&1 | grep -E "ERROR SUMMARY|Error|error|hazard|WARN|launches ok" | head -6
  done
done
