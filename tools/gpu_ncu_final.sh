#!/bin/bash
# final ncu --set full capture of the three hot kernels as the bench runs them (16 frames per sweep launch)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|cost_wide" -s 15 -c 3 -o gpurun_out/prof_r2_final_n16 -f \
    python tools/sweep_probe.py --n 16 --reps 1 --tag ncu > gpurun_out/ncu_full_final.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_final.log
