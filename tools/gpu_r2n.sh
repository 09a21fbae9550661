#!/bin/bash
# Round 2, visit N: runtime rows per band, async batches, the restructured bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgbm_gpu.py tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_r2n.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2n.log
timeout 300 python tools/sweep_probe.py --n 4,8,12,16 --reps 2 --check --tag auto_rows > gpurun_out/probe_r2n.jsonl 2> gpurun_out/probe_r2n.err
python - <<'PY'
import json
for l in open("gpurun_out/probe_r2n.jsonl"):
    d = json.loads(l); print(d["tag"], d["n"], d["ms_per_frame"], d["stage_ms_per_frame"]["aggregate"], d.get("bit_exact_vs_single"))
PY
tail -2 gpurun_out/probe_r2n.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2n.json')); print({k:d[k] for k in ('value','ms_per_step','parity','stage_ms_per_frame')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['ms_per_frame'], d['config']['single_frame_ms'])"; tail -3 gpurun_out/bench_r2n.err
