#!/bin/bash
# Round 2, visit E: the whole GPU suite, smoke, the new bench.py (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2e.log 2>&1
echo "pytest gpu rc=$?"; tail -5 gpurun_out/pytest_gpu_r2e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2e.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2e.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_r2e.json; tail -5 gpurun_out/bench_r2e.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2e.json 2> gpurun_out/bench_ref_r2e.err
echo "bench reference rc=$?"; cut -c1-400 gpurun_out/bench_ref_r2e.json
