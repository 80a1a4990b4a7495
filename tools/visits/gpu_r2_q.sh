#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for c in "c2_doublegauss 0" "c3_asphere 0" "x1_tilted 4000000" "x2_xypoly 4000000" "x6_biconic 4000000"; do timeout 300 python tools/time_kernel.py $c 10; done | tee gpurun_out/timings_q.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --only-extra c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
c=d['config']['extra']['c3']
print("bench c3: gen kernel", c['trace_kernels_ms'], "resident", c['resident_arrays']['trace_kernels_ms'], c['resident_arrays']['roofline']['frac'], "c2 kernel", d['roofline']['kernel_ms'], d['roofline']['frac'])
P
