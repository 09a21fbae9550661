#!/bin/bash
# Round 2, visit K: taller bands (more row warps per SM, shallower rings), n = 8 and 16
mkdir -p gpurun_out; rm -f gpurun_out/probe_r2k.jsonl
for v in r18n6 r17n6 r16n7 r16n6 r20n5; do
  WSG_LIB=$PWD/wass_b200/variants/libwassgpu_$v.so timeout 300 python tools/sweep_probe.py --n 8,16 --reps 2 --check --tag $v >> gpurun_out/probe_r2k.jsonl 2>> gpurun_out/probe_r2k.err
done
timeout 300 python tools/sweep_probe.py --n 16 --reps 2 --tag r14n8 >> gpurun_out/probe_r2k.jsonl 2>> gpurun_out/probe_r2k.err
python - <<'PY'
import json
for l in open("gpurun_out/probe_r2k.jsonl"):
    d = json.loads(l); print(d["tag"], d["n"], d["ms_per_frame"], d["stage_ms_per_frame"]["aggregate"], d.get("bit_exact_vs_single"))
PY
tail -3 gpurun_out/probe_r2k.err
