#!/bin/bash
# Round 2, visit AE: exe tests (async JPEG encoder), workdir wall times with and without the diagnostic images
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_exe.py tests/test_sequence_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/pytest_r2ae.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2ae.log
nproc
timeout 1200 python tools/bench_workdirs.py > gpurun_out/workdirs_r2ae.json 2> gpurun_out/workdirs_r2ae.err
echo "workdirs rc=$?"; cut -c1-1800 gpurun_out/workdirs_r2ae.json; tail -3 gpurun_out/workdirs_r2ae.err
