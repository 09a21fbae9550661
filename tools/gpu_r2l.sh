#!/bin/bash
# parity + stage times after a change to the matcher kernels: tools/gpu_r2l.sh <tag> [batch sizes]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgbm_gpu.py tests/test_fullsize_parity.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_r2l.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2l.log
timeout 300 python tools/sweep_probe.py --n ${2:-8,16} --reps 3 --check --tag ${1:-r2l} > gpurun_out/probe_r2l.jsonl 2> gpurun_out/probe_r2l.err
cut -c1-330 gpurun_out/probe_r2l.jsonl; tail -2 gpurun_out/probe_r2l.err
