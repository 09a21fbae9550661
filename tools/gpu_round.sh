#!/bin/bash
# One full GPU visit: every parity test, smoke, bench (ours + reference arm), config 4, sequence and workdir benches,
# launch list.  Everything is wrapped in `timeout`; logs land in gpurun_out/.
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "bench reference rc=$?"; cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 600 python tools/sweep_probe.py --size 4096x3000x512 --n 1,3 --reps 2 --check --tag config4 > gpurun_out/probe_config4_$TAG.jsonl 2> gpurun_out/probe_config4_$TAG.err
echo "config4 rc=$?"; cut -c1-300 gpurun_out/probe_config4_$TAG.jsonl
timeout 600 python tools/sweep_probe.py --mode 0 --n 16 --reps 2 --check --tag mode_sgbm > gpurun_out/probe_sgbm_$TAG.jsonl 2> gpurun_out/probe_sgbm_$TAG.err
echo "mode sgbm rc=$?"; cut -c1-300 gpurun_out/probe_sgbm_$TAG.jsonl
timeout 600 python tools/bench_sequence.py --frames 32 --mode hh > gpurun_out/sequence_$TAG.json 2> gpurun_out/sequence_$TAG.err
echo "sequence rc=$?"; cut -c1-600 gpurun_out/sequence_$TAG.json
timeout 900 python tools/bench_workdirs.py > gpurun_out/workdirs_$TAG.json 2> gpurun_out/workdirs_$TAG.err
echo "workdirs rc=$?"; cut -c1-900 gpurun_out/workdirs_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
