#!/bin/bash
# closing multi-GPU visit after the plain step / in-order tiles: 8 ranks (weak scaling, device-timed, extras) 
mkdir -p gpurun_out
date
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "concurrent_streams" 2>&1 | tail -3
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 1500 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
date
