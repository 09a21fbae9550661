#!/bin/bash
# Round 2, visit A: the batched sweep kernel -- parity first, then stage times per batch size and per variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_r2a.txt 2>&1
timeout 1200 python -m pytest tests/test_sgbm_gpu.py -x -q -m gpu > gpurun_out/pytest_sgbm_r2a.log 2>&1
echo "pytest sgbm rc=$?"; tail -5 gpurun_out/pytest_sgbm_r2a.log
timeout 900 python -m pytest tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_full_r2a.log 2>&1
echo "pytest fullsize rc=$?"; tail -5 gpurun_out/pytest_full_r2a.log
timeout 300 python tools/sweep_probe.py --n 1,2,4,8 --check --tag r15n5 > gpurun_out/probe_r2a.jsonl 2> gpurun_out/probe_r2a.err
echo "probe rc=$?"; cat gpurun_out/probe_r2a.jsonl; tail -3 gpurun_out/probe_r2a.err
for v in r15n6 r11n8 r11n5 r7n5; do
  WSG_LIB=$PWD/wass_b200/variants/libwassgpu_$v.so timeout 300 python tools/sweep_probe.py --n 1,4,8 --check --tag $v >> gpurun_out/probe_r2a.jsonl 2>> gpurun_out/probe_r2a.err
  echo "probe $v rc=$?"; tail -3 gpurun_out/probe_r2a.jsonl
done
timeout 200 python tools/sweep_probe.py --n 1,4 --mode 0 --check --tag sgbm5path >> gpurun_out/probe_r2a.jsonl 2>> gpurun_out/probe_r2a.err
tail -2 gpurun_out/probe_r2a.jsonl
