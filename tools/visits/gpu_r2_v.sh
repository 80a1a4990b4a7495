#!/bin/bash
# closing evidence after the spot fusion: bench (both arms), launch list, ncu capture of the fused C2 kernel
mkdir -p gpurun_out
date
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
cat > /tmp/prof_fused.py <<'P'
import sys, torch
sys.path.insert(0, '.')
import pyrate_b200 as pb
from pyrate_b200 import configs, engine, lowering
spec = configs.CONFIGS["c2_doublegauss"]
(x0, k0, e0) = configs.config_bundle(spec, spec["bundle"]["rings"])
(s, seq) = configs.build_system(spec, pb.api())
low = lowering.lower(s, seq, configs.DLINE)
dev = torch.device("cuda", 0)
(x0, k0, e0) = engine.device_bundle(x0, k0, e0, dev)
spot = torch.zeros(8, dtype=torch.float64, device=dev)
for _ in range(4):
    engine.trace(low, x0, k0, e0, configs.DLINE, device=dev, spot=(spot, engine.last_surface_origin(low)))
torch.cuda.synchronize()
print("fused launches ok", spot.cpu().numpy()[3])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 2 -c 1 -f -o gpurun_out/prof_r02b_c2spot python /tmp/prof_fused.py > gpurun_out/ncu_c2spot.log 2>&1; tail -1 gpurun_out/ncu_c2spot.log
date
