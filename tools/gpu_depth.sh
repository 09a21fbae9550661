#!/bin/bash
# usage: gpu_depth.sh "ENV=.. --pipeline-depth N" ... : bench (no cpu leg) once per "env-assignments and bench flags" string
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  envs=""; flags=""
  for w in $spec; do case "$w" in *=*) envs="$envs $w";; *) flags="$flags $w";; esac; done
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu $flags > gpurun_out/bench_d$i.json 2> gpurun_out/bench_d$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_d$i.json").read().strip().splitlines()[-1])
    print("[$spec] value %.1f e2e %.1f single_frame_ms %.2f" % (d["value"], d["e2e"]["value"], d["config"].get("single_frame_ms", 0)))
except Exception as e:
    print("[$spec] no json", e); print(open("gpurun_out/bench_d$i.err").read()[-1500:])
PY
done
