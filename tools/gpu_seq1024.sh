#!/bin/bash
# BASELINE configs[4]: 1024-frame synthetic sequence, 2448x2048xD256, N GPUs, NCCL plane reduction, whole frames
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_sequence.py --frames 1024 --mode hh --batch 8 --depth ${2:-3} > gpurun_out/seq1024_${N}gpu.json 2> gpurun_out/seq1024_${N}gpu.err
echo "sequence 1024 frames ${N}gpu rc=$?"; cat gpurun_out/seq1024_${N}gpu.json; tail -3 gpurun_out/seq1024_${N}gpu.err
