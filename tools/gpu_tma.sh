#!/bin/bash
# TMA-staged sweeps (-DWSG_SW_TMA=1, rings 6 deep so that the staging fits) against the register-load build with the same rings
mkdir -p gpurun_out; rm -f gpurun_out/probe_tma.jsonl
V=wass_b200/variants
WSG_LIB=$PWD/$V/libwassgpu_r15n6sw_tma1.so timeout 600 python -m pytest tests/test_sgbm_gpu.py tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_tma.log 2>&1
echo "pytest (TMA variant) rc=$?"; tail -4 gpurun_out/pytest_tma.log
for v in r15n6sw_tma1 r15n6; do
  WSG_LIB=$PWD/$V/libwassgpu_$v.so timeout 300 python tools/sweep_probe.py --n 8,16 --reps 3 --check --tag $v >> gpurun_out/probe_tma.jsonl 2>> gpurun_out/probe_tma.err
done
timeout 300 python tools/sweep_probe.py --n 16 --reps 3 --check --tag default_r15n8 >> gpurun_out/probe_tma.jsonl 2>> gpurun_out/probe_tma.err
WSG_LIB=$PWD/$V/libwassgpu_r15n6sw_tma1.so timeout 300 python tools/sweep_probe.py --size 4096x3000x512 --n 2 --reps 2 --check --tag tma_config4 >> gpurun_out/probe_tma.jsonl 2>> gpurun_out/probe_tma.err
cut -c1-330 gpurun_out/probe_tma.jsonl; tail -3 gpurun_out/probe_tma.err
