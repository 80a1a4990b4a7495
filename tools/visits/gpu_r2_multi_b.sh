#!/bin/bash
# closing multi-GPU visit (gpurun --gpus 8): the headline bench at 8 ranks (weak scaling, device-timed) with
# the BASELINE configs 3-5 ray-sharded (C5: 1e8 rays over 8 GPUs + NCCL spot gather), and at 4 ranks (C4:
# 1e6 rays over 4 GPUs).  The end-to-end legs were measured in the earlier visit (r02_bench_n8.json): they
# are bound by the node's D2H ceiling (r02_pcie_ceiling.md) and are left out here.
mkdir -p gpurun_out
date
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 3000 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
date
timeout 600 $RUN --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e --only-extra c3,c4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 2000 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
date
