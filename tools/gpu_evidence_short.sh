#!/bin/bash
# Short evidence pass: parity, smoke, bench, ncu launch list of the bench command, one
# ncu --set full capture of the headline kernel, kernel timings of the BASELINE configs.
mkdir -p gpurun_out
date
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:trace_real -s 2 -c 1 -f -o gpurun_out/prof_final_c2 python tools/profile_target.py c2_doublegauss 0 4 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
for c in "c1_doublet 1000000" "c2_doublegauss 0" "c3_asphere 0" "c4_anisotropic 1000000" "c5_grin 1000000"; do timeout 100 python tools/time_kernel.py $c 10 | tail -1; done | tee gpurun_out/timings.txt
date
