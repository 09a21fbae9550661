#!/bin/bash
# usage: gpu_env.sh "ENV=.. ENV2=.." ...   : bench (no cpu leg) once per environment string
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_env$i.json 2> gpurun_out/bench_env$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_env$i.json").read().strip().splitlines()[-1])
    print("[$envs] value %.1f e2e %.1f single_frame_ms %.2f stages %s" % (d["value"], d["e2e"]["value"], d["config"].get("single_frame_ms", 0), {k: round(v, 3) for k, v in d["stage_ms_per_frame"].items()}))
except Exception as e:
    print("[$envs] no json", e); print(open("gpurun_out/bench_env$i.err").read()[-1500:])
PY
done
