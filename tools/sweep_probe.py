#!/usr/bin/env python
"""Times the dense matcher's stages for batches of n frames (device-resident inputs, CUDA events via wsg_profile_*).

    python tools/sweep_probe.py [--n 1,2,4,8] [--reps 3] [--mode 1] [--size 2448x2048x256] [--check]

Prints one JSON line per batch size: per-frame stage times, the aggregation's share, and (with --check) whether every
frame of the batch equals the single-frame result bit for bit.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--size", default="2448x2048x256")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--impl", type=int, default=2)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    from wass_b200 import capi, synth
    W, H, D = (int(v) for v in a.size.split("x"))
    win = 13
    p = dict(minDisparity=1, numDisparities=D, blockSize=win, P1=2 * win * win, P2=64 * win * win, disp12MaxDiff=-1,
             preFilterCap=60, uniquenessRatio=1, speckleWindowSize=-70, speckleRange=16, mode=a.mode)
    ns = [int(v) for v in a.n.split(",")]
    nmax = max(ns)
    frames = [synth.pad_for_sgbm(*synth.make_pair(W, H, D, seed=s)[:2], D) for s in range(min(nmax, 4))]
    Hh, Wp = frames[0][0].shape
    i1 = torch.from_numpy(np.stack([frames[f % len(frames)][0] for f in range(nmax)])).cuda()
    i2 = torch.from_numpy(np.stack([frames[f % len(frames)][1] for f in range(nmax)])).cuda()
    out = torch.empty((nmax, Hh, Wp), dtype=torch.int16, device="cuda")
    h = capi.Handle(0)
    h.sgbm_set_impl(a.impl)
    st = torch.cuda.Stream()
    h.set_stream(st.cuda_stream)
    ref = None
    if a.check:
        ref = []
        for f in range(len(frames)):
            h.sgbm_compute_batch_device(1, i1[f].data_ptr(), i2[f].data_ptr(), Hh * Wp, Hh, Wp, Wp, p, out[0].data_ptr())
            h.synchronize()
            ref.append(out[0].clone())
    for n in ns:
        def run():
            h.sgbm_compute_batch_device(n, i1.data_ptr(), i2.data_ptr(), Hh * Wp, Hh, Wp, Wp, p, out.data_ptr())
        run(); h.synchronize()
        stats = h.sgbm_stats()      # raises on a sweep error
        h.profile_enable(True); h.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.reps):
            run()
        e1.record(st)
        h.synchronize()
        total = e0.elapsed_time(e1) / a.reps / n
        prof = h.profile_get()
        h.profile_enable(False)
        line = {"tag": a.tag, "size": a.size, "mode": a.mode, "impl": a.impl, "n": n, "ms_per_frame": round(total, 3),
                "Mdisp_s": round(W * H / total / 1e3, 1),
                "stage_ms_per_frame": {k: round(v[0] / a.reps / n, 3) for k, v in prof.items() if v[1]},
                "max_cost": stats["max_cost"]}
        V = stats["volume_bytes"]
        agg = prof["aggregate"][0] / a.reps / n
        line["agg_gbs"] = round(4 * V / agg / 1e6, 1)
        if a.check:
            line["bit_exact_vs_single"] = all(bool(torch.equal(out[f], ref[f % len(frames)])) for f in range(n))
        print(json.dumps(line), flush=True)
    h.close()


if __name__ == "__main__":
    main()
