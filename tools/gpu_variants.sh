#!/bin/bash
timeout 300 python tools/time_kernel.py c2_doublegauss 0 20
timeout 300 python tools/time_kernel.py c2_doublegauss 0 10 1
timeout 300 python tools/time_kernel.py c3_asphere 0 10
timeout 300 python tools/time_kernel.py x2_xypoly 4000000 10
timeout 300 python tools/time_kernel.py x6_biconic 4000000 10
timeout 300 python tools/time_kernel.py c5_grin 1000000 5
