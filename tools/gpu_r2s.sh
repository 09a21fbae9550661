#!/bin/bash
# Round 2, visit S: ncu --set full of the cost kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cost_wide" -s 2 -c 2 -o gpurun_out/prof_${1:-r2s}_cost -f \
    python tools/sweep_probe.py --n 4 --reps 1 --tag ncu > gpurun_out/ncu_full_r2s.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_r2s.log
