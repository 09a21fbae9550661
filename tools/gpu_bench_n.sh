#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench ${N}gpu rc=$?"; cut -c1-400 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
