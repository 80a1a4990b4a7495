#!/bin/bash
timeout 200 python tools/time_kernel.py x10_zernike_general 4000000 5
timeout 200 python tools/time_kernel.py x2_xypoly 4000000 5
