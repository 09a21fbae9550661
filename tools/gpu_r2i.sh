#!/bin/bash
# Round 2, visit I: ncu --set full of the current sweeps (R=14, NS=8) and the cost kernel at n=4; workdir wall times again
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|cost_wide" -s 6 -c 6 -o gpurun_out/prof_r2i_n4 -f \
    python tools/sweep_probe.py --n 4 --reps 1 --tag ncu > gpurun_out/ncu_full_r2i.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_r2i.log
timeout 600 python tools/bench_workdirs.py --frames 16 --parallel 4 > gpurun_out/workdirs_r2i.json 2> gpurun_out/workdirs_r2i.err
echo "workdirs rc=$?"; cat gpurun_out/workdirs_r2i.json; tail -3 gpurun_out/workdirs_r2i.err
