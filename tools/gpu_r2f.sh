#!/bin/bash
# Round 2, visit F: the rest of the GPU suite; config 4 (4096x3000x512) stage times; the whole-frame sequence on one GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2f.log 2>&1
echo "pytest gpu rc=$?"; tail -8 gpurun_out/pytest_gpu_r2f.log
timeout 600 python tools/sweep_probe.py --size 4096x3000x512 --n 1,2,3 --reps 2 --check --tag config4 > gpurun_out/probe_config4_r2f.jsonl 2> gpurun_out/probe_config4_r2f.err
echo "config4 rc=$?"; cut -c1-400 gpurun_out/probe_config4_r2f.jsonl; tail -3 gpurun_out/probe_config4_r2f.err
timeout 600 python tools/bench_sequence.py --frames 32 --mode hh --batch 8 --depth 2 > gpurun_out/seq_n1_r2f.json 2> gpurun_out/seq_n1_r2f.err
echo "sequence rc=$?"; cat gpurun_out/seq_n1_r2f.json; tail -3 gpurun_out/seq_n1_r2f.err
