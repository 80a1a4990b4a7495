#!/bin/bash
PYR_LEAN_VARIANT=0 timeout 200 python tools/compare_variants.py save
for v in 70 61 62 63; do PYR_LEAN_VARIANT=$v timeout 200 python tools/compare_variants.py check | tail -1; done
for v in 0 70 61 62 63; do PYR_LEAN_VARIANT=$v timeout 200 python tools/time_kernel.py c2_doublegauss 0 20; done
