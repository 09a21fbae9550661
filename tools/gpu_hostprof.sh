#!/bin/bash
# where the host time of `wass_stereo --batch` goes (WASS_HOST_PROFILE=1), 16 workdirs of 2448x2048x256
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/hostprof.log 2>&1
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from wass_b200 import synth, workdir
W, H, D, n = 2448, 2048, 256, 16
c = synth.make_calibration(W, H)
td = tempfile.mkdtemp()
pairs = [synth.make_pair(W, H, D, seed=s, d0=16.0) for s in range(4)]
for dbg in (False, True):
    cfg = os.path.join(td, "cfg%d.txt" % dbg); workdir.write_config(cfg, MAX_DISPARITY=D, RANDOM_SEED=1, SGM_FULL_8PATH=True, SAVE_DEBUG_IMAGES=dbg)
    wds = []
    for i in range(n):
        r, l, _ = pairs[i % 4]
        wd = os.path.join(td, "d%d_%06d_wd" % (dbg, i)); workdir.write_workdir(wd, l, r, c["K0"], c["K1"], c["R"], c["T"]); wds.append(wd)
    t0 = time.perf_counter()
    p = subprocess.run(["wass_b200/bin/wass_stereo", "--batch", "--batch-size", "8", cfg] + wds, capture_output=True, text=True,
                       env=dict(os.environ, WASS_HOST_PROFILE="1"))
    print("debug images", dbg, "rc", p.returncode, "wall %.2f s" % (time.perf_counter() - t0))
    print(p.stdout.strip().splitlines()[-2])
    print(p.stderr)
PY
cat gpurun_out/hostprof.log
