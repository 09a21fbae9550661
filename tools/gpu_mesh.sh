#!/bin/bash
# mesh-stage kernels: parity tests (per-test timeout), per-kernel times of a short whole-frame sequence (ncu launch list),
# sequence throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_sequence_gpu.py tests/test_host_exe.py -x -q -m gpu --timeout 120 > gpurun_out/pytest_mesh.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mesh.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_seq.csv \
    python tools/bench_sequence.py --frames 2 --mode hh --depth 1 --batch 2 > gpurun_out/ncu_seq.log 2>&1
echo "ncu rc=$?"
timeout 300 python tools/bench_sequence.py --frames 96 --mode hh --depth 3 --batch 8 > gpurun_out/seq_mesh.json 2> gpurun_out/seq_mesh.err
cut -c1-700 gpurun_out/seq_mesh.json
