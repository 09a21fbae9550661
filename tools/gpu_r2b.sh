#!/bin/bash
# Round 2, visit B: ncu --set full (with source) of the two batched sweep launches, n = 4 frames, variant r15n6
mkdir -p gpurun_out
export WSG_LIB=$PWD/wass_b200/variants/libwassgpu_r15n6.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel" -s 2 -c 2 -o gpurun_out/prof_r2b_sweeps_n4 -f \
    python tools/sweep_probe.py --n 4 --reps 1 --tag ncu > gpurun_out/ncu_full_r2b.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_r2b.log
ls -la gpurun_out/*.ncu-rep
