#!/usr/bin/env python
"""Wall time per workdir of the drop-in executable, files in -> files out (PNG decode, rectification, dense matcher,
triangulation, plane, mesh_cam.xyzC written to disk), the contract of cli/wasscli/wasscli.py:326-346:

  (i)  one wass_stereo process per frame, P at a time (what the reference's drivers and wass_b200/launcher.py do): every
       frame pays process start, CUDA context creation and the device arena;
  (ii) `wass_stereo --batch`: one process, one warm arena, batched matcher runs, PNG decode overlapped.

    python tools/bench_workdirs.py [--frames 16] [--size 2448x2048x256] [--parallel 4] [--mode hh]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wass_b200", "bin", "wass_stereo")


def main():
    from wass_b200 import synth, workdir
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--size", default="2448x2048x256")
    ap.add_argument("--parallel", type=int, default=4, help="processes at a time in the one-process-per-frame arm (wasscli default 4)")
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--mode", default="hh", choices=["sgbm", "hh"])
    a = ap.parse_args()
    W, H, D = (int(v) for v in a.size.split("x"))
    c = synth.make_calibration(W, H)
    pairs = [synth.make_pair(W, H, D, seed=s, d0=16.0) for s in range(min(a.frames, 4))]
    out = {"workload": "%dx%d D=%d %s, %d workdirs, files in -> files out" % (W, H, D, a.mode, a.frames)}
    # the reference always writes its diagnostic JPEGs (six images, ~60 MP to encode per frame); SAVE_DEBUG_IMAGES=false is
    # this build's switch for production runs
    for debug_images in (True, False):
        with tempfile.TemporaryDirectory() as td:
            cfg = os.path.join(td, "stereo_config.txt")
            workdir.write_config(cfg, MAX_DISPARITY=D, RANDOM_SEED=1, SGM_FULL_8PATH=(a.mode == "hh"), SAVE_DEBUG_IMAGES=debug_images)

            def make(tag):
                wds = []
                for i in range(a.frames):
                    right, left, _ = pairs[i % len(pairs)]
                    wd = os.path.join(td, "%s_%06d_wd" % (tag, i))
                    workdir.write_workdir(wd, left, right, c["K0"], c["K1"], c["R"], c["T"])
                    wds.append(wd)
                return wds
            w1, w2 = make("p"), make("b")
            subprocess.run([EXE, cfg, w1[0]], capture_output=True)          # page the binaries in
            t0 = time.perf_counter()
            with ThreadPoolExecutor(a.parallel) as ex:
                rcs = list(ex.map(lambda wd: subprocess.run([EXE, cfg, wd], capture_output=True).returncode, w1))
            t_proc = time.perf_counter() - t0
            t0 = time.perf_counter()
            r = subprocess.run([EXE, "--batch", "--batch-size", str(a.batch_size), "--planes-out", os.path.join(td, "planes.txt"), cfg] + w2,
                               capture_output=True, text=True)
            t_batch = time.perf_counter() - t0
            log = open(os.path.join(w2[-1], "wass_stereo_log.txt")).read()
            table = [l.split("] ", 1)[1] for l in log.splitlines() if l.startswith("wass_stereo [info ] |") and "Task" not in l]
            same = all(open(os.path.join(x, "mesh_cam.xyzC"), "rb").read() == open(os.path.join(y, "mesh_cam.xyzC"), "rb").read()
                       for x, y in zip(w1, w2))
            out["debug_images_%s" % ("on" if debug_images else "off")] = {
                "process_per_frame": {"parallel": a.parallel, "ms_per_workdir": t_proc / a.frames * 1e3, "failed": sum(1 for x in rcs if x)},
                "batch_process": {"batch_size": a.batch_size, "ms_per_workdir": t_batch / a.frames * 1e3, "rc": r.returncode,
                                  "tail": r.stdout.strip().splitlines()[-2:]},
                "identical_xyzC": same, "host_timer_of_the_last_workdir": table}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
