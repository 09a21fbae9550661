#!/bin/bash
# N GPUs of one box (gpurun --gpus N -- bash tools/gpu_multi.sh N): bench.py, the whole-frame sequence bench (BASELINE config 5 at N = 8)
# and the C++ batch host with one rank per GPU -- everything that uses NCCL through the C ABI.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench ${N}gpu rc=$?"; cut -c1-600 gpurun_out/bench_${N}gpu.json; grep -o '"plane_reduction.*' gpurun_out/bench_${N}gpu.json | cut -c1-500; tail -4 gpurun_out/bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_sequence.py --frames $((32*N)) --mode hh --batch 8 --depth 2 > gpurun_out/seq_${N}gpu.json 2> gpurun_out/seq_${N}gpu.err
echo "sequence ${N}gpu rc=$?"; cat gpurun_out/seq_${N}gpu.json; tail -4 gpurun_out/seq_${N}gpu.err
# the C++ batch host on two ranks (one process per GPU, NCCL id through a file)
NRANKS=$N python - <<'PY' > gpurun_out/exe_${N}rank.log 2>&1
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.getcwd())
from wass_b200 import synth, workdir
NR = int(os.environ["NRANKS"])
W, H, D, n = 640, 480, 64, 4 * NR
c = synth.make_calibration(W, H)
td = tempfile.mkdtemp()
cfg = os.path.join(td, "cfg.txt"); workdir.write_config(cfg, MAX_DISPARITY=D, RANDOM_SEED=3, PLANE_RANSAC_ROUNDS=60)
wds = []
for i in range(n):
    r, l, _ = synth.make_pair(W, H, D, seed=i, d0=8.0 + 0.5 * i)
    wd = os.path.join(td, "%06d_wd" % i); workdir.write_workdir(wd, l, r, c["K0"], c["K1"], c["R"], c["T"]); wds.append(wd)
exe = "wass_b200/bin/wass_stereo"
ps = [subprocess.Popen([exe, "--batch", "--batch-size", "2", "--ranks", str(NR), "--rank", str(r), "--nccl-id-file", os.path.join(td, "id"),
                        "--planes-out", os.path.join(td, "planes.txt"), cfg] + wds, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, WASS_GPU_DEVICE=str(r))) for r in range(NR)]
outs = [p.communicate(timeout=300)[0] for p in ps]
print([p.returncode for p in ps])
for o in outs: print("\n".join(o.strip().splitlines()[-3:]))
rows = np.loadtxt(os.path.join(td, "planes.txt"))
mine = np.array([[float(v) for v in open(os.path.join(w, "plane.txt")).read().split()] for w in wds])
print("planes file in frame order:", bool(np.allclose(rows, mine, atol=1e-15)), "mean", np.nanmean(mine, axis=0))
PY
echo "exe $N ranks rc=$?"; tail -9 gpurun_out/exe_${N}rank.log
