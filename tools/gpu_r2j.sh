#!/bin/bash
# Round 2, visit J: band timelines (fill / drain of the batched wavefronts)
mkdir -p gpurun_out; rm -f gpurun_out/timeline_r2j.jsonl
for n in 1 4 8 16; do
  timeout 300 python tools/band_timeline.py --n $n >> gpurun_out/timeline_r2j.jsonl 2>> gpurun_out/timeline_r2j.err
done
cat gpurun_out/timeline_r2j.jsonl; tail -3 gpurun_out/timeline_r2j.err
