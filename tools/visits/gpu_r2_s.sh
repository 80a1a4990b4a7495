#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for c in "c2_doublegauss 0" "c2_doublegauss 0 10 1" "c1_doublet 4000000" "c3_asphere 0" "x14_dispersive_dg 4000000"; do timeout 300 python tools/time_kernel.py $c 2>&1 | tail -1; done | tee gpurun_out/timings_s.txt
timeout 300 python tools/time_gen.py c2_doublegauss 0 10 2>&1 | head -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; tail -2 gpurun_out/bench_s.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_s.json').read().strip().splitlines()[-1])
print("bench: value %.4e ms %.4f kernel %.4f frac %.4f | gen ms %.4f kernel %.4f | e2e %.4e (%.2f ms) host %.2f metric %.3f" % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['generated']['ms_per_step'], d['config']['generated']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_buffers']['ms_per_step'], d['e2e']['metric_only']['ms_per_step']))
P
timeout 120 python tools/time_small.py | grep -i merit
