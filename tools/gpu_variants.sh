#!/bin/bash
timeout 300 python tools/time_kernel.py c2_doublegauss 0 20
timeout 300 python tools/time_kernel.py c1_doublet 1000000 10
timeout 300 python tools/time_kernel.py x1_tilted 4000000 10
timeout 300 python tools/time_kernel.py c3_asphere 0 10
timeout 300 python tools/time_kernel.py c2_doublegauss 0 10 1
