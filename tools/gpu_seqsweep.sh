#!/bin/bash
# whole-frame sequence throughput on one GPU for several (handles, frames per batch)
mkdir -p gpurun_out; rm -f gpurun_out/seqsweep.jsonl
for cfg in "2 8" "3 4" "3 8" "2 12" "4 4" "1 16"; do
  set -- $cfg
  timeout 300 python tools/bench_sequence.py --frames 96 --mode hh --depth $1 --batch $2 >> gpurun_out/seqsweep.jsonl 2>> gpurun_out/seqsweep.err
done
python - <<'PY'
import json
for l in open("gpurun_out/seqsweep.jsonl"):
    d = json.loads(l); print(d["handles_per_gpu"], d["frames_per_batch"], round(d["ms_per_frame_per_gpu"], 2), round(d["frames_per_s"], 1))
PY
tail -3 gpurun_out/seqsweep.err
