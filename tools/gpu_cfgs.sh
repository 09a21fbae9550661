#!/bin/bash
mkdir -p gpurun_out
for cfg in "$@"; do
  WSG_SWEEP_CFG=$cfg timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_cfg$cfg.json 2> gpurun_out/bench_cfg$cfg.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_cfg$cfg.json").read().strip().splitlines()[-1])
    print("cfg $cfg value %.1f e2e %.1f single_frame_ms %.2f stages %s" % (d["value"], d["e2e"]["value"], d["config"].get("single_frame_ms", 0), d["stage_ms_per_frame"]))
except Exception as e:
    print("cfg $cfg no json", e); print(open("gpurun_out/bench_cfg$cfg.err").read()[-1500:])
PY
done
