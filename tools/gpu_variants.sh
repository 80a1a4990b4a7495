#!/bin/bash
timeout 200 python tools/time_small.py c2_doublegauss 18 2>&1 | tail -7
