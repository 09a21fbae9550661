#!/bin/bash
# Round 2, visit Q: exe tests (diagnostic JPEGs), DRAM traffic of the kernels of one batch of 16 frames, launch list of the
# bench command, the bench itself.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_exe.py -x -q -m gpu > gpurun_out/pytest_r2q.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2q.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"sweep_kernel|cost_wide" -o gpurun_out/traffic_r2q -f python tools/sweep_probe.py --n 16 --reps 1 --tag ncu > gpurun_out/traffic_r2q.log 2>&1
echo "traffic rc=$?"; tail -2 gpurun_out/traffic_r2q.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2q.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_r2q.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2q.json; tail -3 gpurun_out/bench_r2q.err
