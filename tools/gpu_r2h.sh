#!/bin/bash
# Round 2, visit H: cost kernel without the pre-negated tables (parity + time), workdir wall times, 1-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgbm_gpu.py tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_r2h.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2h.log
timeout 300 python tools/sweep_probe.py --n 8 --reps 3 --tag cost6tab > gpurun_out/probe_r2h.jsonl 2> gpurun_out/probe_r2h.err
cut -c1-330 gpurun_out/probe_r2h.jsonl; tail -2 gpurun_out/probe_r2h.err
timeout 900 python tools/bench_workdirs.py --frames 16 --parallel 4 > gpurun_out/workdirs_r2h.json 2> gpurun_out/workdirs_r2h.err
echo "workdirs rc=$?"; cat gpurun_out/workdirs_r2h.json; tail -3 gpurun_out/workdirs_r2h.err
