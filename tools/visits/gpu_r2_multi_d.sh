#!/bin/bash
# closing build at 4 and 2 ranks (device-timed; extras at 4 ranks: C4 1e6 rays over 4 GPUs)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e --only-extra c3,c4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 600 gpurun_out/bench_n4.json; tail -2 gpurun_out/bench_n4.err
timeout 600 $RUN --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-extras > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
