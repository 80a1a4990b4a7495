#!/bin/bash
# round 2, multi-GPU visit (gpurun --gpus 8): node PCIe ceiling, the headline bench at 8 ranks
# (weak scaling; extras = BASELINE configs 3-5 ray-sharded: C5 1e8 rays over 8 GPUs + NCCL spot
# gather) and at 4 ranks (C4: 1e6 rays over 4 GPUs)
mkdir -p gpurun_out
date
nvidia-smi -L | tee gpurun_out/gpus.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" | tee gpurun_out/lscpu.txt
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $RUN --nproc-per-node 8 --master-port 29511 tools/pcie_ceiling.py > gpurun_out/pcie_ceiling_n8.json 2> gpurun_out/pcie_n8.err; tail -c 1500 gpurun_out/pcie_ceiling_n8.json
timeout 300 $RUN --nproc-per-node 8 --master-port 29512 tools/pcie_ceiling.py --no-bind > gpurun_out/pcie_ceiling_n8_nobind.json 2> gpurun_out/pcie_n8_nobind.err; tail -c 600 gpurun_out/pcie_ceiling_n8_nobind.json
timeout 120 python tools/pcie_ceiling.py > gpurun_out/pcie_ceiling_n1.json 2>/dev/null; tail -c 400 gpurun_out/pcie_ceiling_n1.json
date
timeout 900 $RUN --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 6000 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
date
timeout 600 $RUN --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 3000 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
date
timeout 600 $RUN --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json
date
