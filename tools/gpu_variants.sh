#!/bin/bash
for v in 0 3 31 32 33; do PYR_LEAN_VARIANT=$v timeout 300 python tools/time_kernel.py c2_doublegauss 0 20; done
