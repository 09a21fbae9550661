#!/bin/bash
# Short GPU visit: matcher parity, bench of the default implementation, optional ncu of a kernel regex.
TAG=${1:-q}; KREGEX=${2:-}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgbm_gpu.py -x -q -m gpu > gpurun_out/pytest_sgbm_$TAG.log 2>&1
echo "pytest sgbm rc=$?"; tail -3 gpurun_out/pytest_sgbm_$TAG.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f frac %.3f single_frame_ms %.2f stages %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("single_frame_ms", 0), d["stage_ms_per_frame"]))
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
if [ -n "$KREGEX" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 2 -c 2 -o gpurun_out/prof_$TAG -f \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu rc=$?"
fi
