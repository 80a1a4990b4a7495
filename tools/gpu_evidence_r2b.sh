#!/bin/bash
# Round-2 closing evidence (one B200): parity, smoke, bench (both arms), ncu launch list of the bench
# command, ncu --set full captures of every kernel family, timings, sanitizer passes.
mkdir -p gpurun_out
date
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
CAPTURES=${CAPTURES:-"c2_doublegauss:0:trace_real:2:c2:mem c2_doublegauss:0:trace_real:2:c2gen:gen c3_asphere:0:trace_real:1:c3:mem c4_anisotropic:1000000:trace_complex:1:c4:mem c5_grin:1000000:trace_real:1:c5:mem"}
for cfg in $CAPTURES; do
  cfg=${cfg//:/ }
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/prof_r02b_$5 python tools/profile_target.py $1 $2 4 $6 > gpurun_out/ncu_$5.log 2>&1; tail -1 gpurun_out/ncu_$5.log
done
for c in "c1_doublet 1000000" "c2_doublegauss 0" "c3_asphere 0" "c4_anisotropic 1000000" "c5_grin 1000000" "x1_tilted 4000000" "x2_xypoly 4000000" "x4_biaxial 1000000" "x6_biconic 4000000" "x16_cylinder 4000000"; do timeout 300 python tools/time_kernel.py $c 10; done | tee gpurun_out/timings.txt
for c in "c2_doublegauss 0" "c3_asphere 0"; do timeout 300 python tools/time_gen.py $c 10 2>&1 | head -5; done | tee gpurun_out/time_gen.txt
timeout 120 python tools/time_small.py | tee gpurun_out/small_latency.txt
timeout 120 python tools/membw.py | tee gpurun_out/membw.txt
for run in "memcheck c2_doublegauss 20000" "memcheck c3_asphere 9000" "memcheck c5_grin 3000" "memcheck c4_anisotropic 3000" "memcheck x1_tilted 7001" "racecheck c2_doublegauss 20000" "racecheck c3_asphere 9000"; do
  set -- $run
  echo "== $1 $2 $3"
  timeout 280 compute-sanitizer --tool $1 --print-limit 5 python tools/profile_target.py $2 $3 1 2>&1 | grep -E "ERROR SUMMARY|Error|error|hazard|WARN|launches ok" | head -6
done | tee gpurun_out/sanitizer.txt
date
