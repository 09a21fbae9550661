#!/bin/bash
# compute-sanitizer over the kernels of round 2: memcheck on smoke(), small batched matcher runs and the mesh stages;
# racecheck on the cost kernel (three concurrent stages sharing shared-memory arrays) and the tiled component labelling
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck smoke: $(grep -E 'ERROR SUMMARY' gpurun_out/memcheck_smoke.log | tail -1)"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_sgbm_gpu.py -x -q -m gpu -k "batch or async or rows_per_band or edge" --timeout 1200 > gpurun_out/memcheck_sgbm.log 2>&1
echo "memcheck sgbm: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/memcheck_sgbm.log | tail -2 | tr '\n' ' ')"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu -k "not full_frame" --timeout 1200 > gpurun_out/memcheck_pipeline.log 2>&1
echo "memcheck pipeline: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/memcheck_pipeline.log | tail -2 | tr '\n' ' ')"
timeout 900 $CS --tool racecheck --kernel-regex kns=cost_wide python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_cost.log 2>&1
echo "racecheck cost: $(grep -E 'RACECHECK SUMMARY' gpurun_out/racecheck_cost.log | tail -1)"
timeout 900 $CS --tool racecheck --kernel-regex kns=ccl_local python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu -k "mesh_ops or tie_rule" --timeout 800 > gpurun_out/racecheck_ccl.log 2>&1
echo "racecheck ccl_local: $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/racecheck_ccl.log | tail -2 | tr '\n' ' ')"
