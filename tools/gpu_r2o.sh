#!/bin/bash
# Round 2, visit O: ncu --set full of the sweeps (register prefetch, R=15, NS=8), n = 4 -> rows picked by the model
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel" -s 2 -c 2 -o gpurun_out/prof_r2o_n4 -f \
    python tools/sweep_probe.py --n 4 --reps 1 --tag ncu > gpurun_out/ncu_full_r2o.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_r2o.log
