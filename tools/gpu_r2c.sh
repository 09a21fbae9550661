#!/bin/bash
# Round 2, visit C: unrolled sweep rows + staggered rows + deeper rings: parity, then stage times per variant
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sgbm_gpu.py tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_r2c.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2c.log
rm -f gpurun_out/probe_r2c.jsonl
timeout 300 python tools/sweep_probe.py --n 1,4,8 --check --tag r14n8 >> gpurun_out/probe_r2c.jsonl 2>> gpurun_out/probe_r2c.err
for v in r14n8sw_stagger0 r14n8sw_stagger1 r15n7 r13n8 r15n6; do
  WSG_LIB=$PWD/wass_b200/variants/libwassgpu_$v.so timeout 300 python tools/sweep_probe.py --n 4,8 --check --tag $v >> gpurun_out/probe_r2c.jsonl 2>> gpurun_out/probe_r2c.err
done
timeout 200 python tools/sweep_probe.py --n 4 --mode 0 --check --tag sgbm5path >> gpurun_out/probe_r2c.jsonl 2>> gpurun_out/probe_r2c.err
cut -c1-330 gpurun_out/probe_r2c.jsonl; tail -3 gpurun_out/probe_r2c.err
