#!/bin/bash
# One GPU visit: parity tests, smoke, bench (ours + reference arm + the other aggregation implementations), ncu evidence.
# Everything is wrapped in `timeout`; logs land in gpurun_out/.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "bench reference rc=$?"; cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
for impl in 1 0; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --agg-impl $impl --pipeline-depth 1 > gpurun_out/bench_${TAG}_impl$impl.json 2> gpurun_out/bench_${TAG}_impl$impl.err
  echo "bench impl $impl rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --pipeline-depth 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|cost_wide" -s 3 -c 3 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --pipeline-depth 1 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
