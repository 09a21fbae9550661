#!/bin/bash
# One GPU visit: parity tests first (bounded), then bench per aggregation implementation, then ncu evidence.
# Everything is wrapped in `timeout`; logs land in gpurun_out/.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_sgbm_gpu.py -x -q -m gpu > gpurun_out/pytest_sgbm_$TAG.log 2>&1
echo "pytest sgbm rc=$?" | tee -a gpurun_out/summary_$TAG.txt
tail -5 gpurun_out/pytest_sgbm_$TAG.log
for impl in 2 1 0; do
  timeout 300 python bench.py --steps 5 --warmup 3 --agg-impl $impl $( [ $impl != 2 ] && echo --no-cpu ) > gpurun_out/bench_${TAG}_impl$impl.json 2> gpurun_out/bench_${TAG}_impl$impl.err
  echo "bench impl $impl rc=$?" | tee -a gpurun_out/summary_$TAG.txt
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_impl$impl.json").read().strip().splitlines()[-1])
    print("impl $impl value %.1f e2e %.1f frac %.3f stages %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["stage_ms_per_frame"]))
except Exception as e:
    print("bench impl $impl: no json", e)
PY
done
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_sgbm_gpu.py > gpurun_out/pytest_rest_$TAG.log 2>&1
echo "pytest rest rc=$?" | tee -a gpurun_out/summary_$TAG.txt
tail -3 gpurun_out/pytest_rest_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 2 -o gpurun_out/prof_sweep_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_sweep_$TAG.log 2>&1
echo "ncu sweep rc=$?" | tee -a gpurun_out/summary_$TAG.txt
