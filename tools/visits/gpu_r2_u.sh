#!/bin/bash
# fused spot sums (POLICY bit 32): parity + step timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; tail -2 gpurun_out/bench_u.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
print("bench: value %.4e ms %.4f kernel %.4f frac %.4f | gen ms %.4f kernel %.4f | e2e %.4e (%.2f ms) metric %.3f rms %.12f" % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['generated']['ms_per_step'], d['config']['generated']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['metric_only']['ms_per_step'], d['config']['spot_rms']))
P
timeout 120 python tools/time_small.py | grep -i "merit\|rms"
