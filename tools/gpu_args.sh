#!/bin/bash
# usage: gpu_args.sh "ENV=.. -- bench args" ...
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  envs="${spec%%--*}"; args="${spec#*--}"
  env $envs timeout 300 python bench.py --no-cpu $args > gpurun_out/bench_arg$i.json 2> gpurun_out/bench_arg$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_arg$i.json").read().strip().splitlines()[-1])
    print("[$spec] value %.1f (%.2f ms/frame) e2e %.1f single_frame_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("single_frame_ms", 0)))
except Exception as e:
    print("[$spec] no json", e); print(open("gpurun_out/bench_arg$i.err").read()[-1500:])
PY
done
