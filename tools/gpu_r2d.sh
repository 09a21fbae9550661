#!/bin/bash
# Round 2, visit D: how long a waiting row polls before it sleeps (WSG_SWEEP_EAGER), n = 8
mkdir -p gpurun_out; rm -f gpurun_out/probe_r2d.jsonl
for e in 0 1 2 4 16 64 256; do
  WSG_SWEEP_EAGER=$e timeout 200 python tools/sweep_probe.py --n 8 --reps 2 --tag eager$e >> gpurun_out/probe_r2d.jsonl 2>> gpurun_out/probe_r2d.err
done
cut -c1-300 gpurun_out/probe_r2d.jsonl; tail -3 gpurun_out/probe_r2d.err
