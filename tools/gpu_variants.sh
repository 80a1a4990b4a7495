#!/bin/bash
for v in 0 1 3 4 11 12; do PYR_LEAN_VARIANT=$v python tools/time_kernel.py c2_doublegauss 0 20; done
python tools/time_kernel.py c2_doublegauss 0 10 1
python tools/time_kernel.py c1_doublet 1000000 10
python tools/time_kernel.py x1_tilted 4000000 10
python tools/time_kernel.py c3_asphere 0 10
python tools/time_kernel.py c5_grin 1000000 5
python tools/time_kernel.py c4_anisotropic 1000000 5
