#!/usr/bin/env python
"""End-to-end throughput of the whole per-frame hot path (BASELINE configs[4] shape: 2448x2048, D=256, stereo +
triangulation + PovMesh + .xyzC in memory, NaN-aware plane all-reduce at the end), in process, through the C ABI.

    python tools/bench_sequence.py [--frames F] [--mode sgbm|hh]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sequence.py --frames F

Host buffers in, host buffers out (every call copies its inputs to the device and its results back); F distinct synthetic
calibrated frames are cycled.  Prints one JSON line on rank 0.  A secondary figure: bench.py stays the headline benchmark.
"""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from wass_b200 import capi, sequence, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16, help="frames of the whole sequence")
    ap.add_argument("--distinct", type=int, default=2, help="distinct synthetic frames generated per rank (cycled)")
    ap.add_argument("--mode", default="sgbm", choices=["sgbm", "hh"])
    ap.add_argument("--depth", type=int, default=3, help="frames in flight per GPU (one handle, stream and host thread each)")
    ap.add_argument("--sweep-workers", type=int, default=-1, help="SMs per fused sweep (0 = all; -1 = half the SMs when 3+ frames are in flight)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, D = 2448, 2048, 256
    c = synth.make_calibration(W, H)
    calib = sequence.rectified_calib(c["K0"], c["K1"], c["R"], c["T"], W, H)
    pool = []
    for s in range(a.distinct):
        right, left, _ = synth.make_pair(W, H, D, seed=1000 * rank + s, d0=16.0)
        pool.append((left, right))
    frames = [pool[i % a.distinct] for i in range(a.frames)]
    dense = capi.dense_params(MAX_DISPARITY=D, mode=capi.MODE_HH if a.mode == "hh" else capi.MODE_SGBM)
    # warm-up (arena allocation, first launches) on the handle the timed run uses
    h = [capi.Handle(local) for _ in range(a.depth)]
    if a.sweep_workers < 0:
        a.sweep_workers = torch.cuda.get_device_properties(local).multi_processor_count // 2 if a.depth >= 3 else 0
    for x in h:
        x.sgbm_set_sweep_workers(a.sweep_workers)
    xyzc_out = [torch.empty(148 + 6 * W * H, dtype=torch.uint8).pin_memory().numpy() for _ in range(a.depth)]   # reusable pinned destinations
    sequence.run_sequence(frames[: 2 * a.depth * world], calib, dense, device=local, rank=rank, world=world, dist=None, handle=h, xyzc_out=xyzc_out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean, planes, res = sequence.run_sequence(frames, calib, dense, device=local, rank=rank, world=world,
                                              dist=dist if world > 1 else None, handle=h, xyzc_out=xyzc_out)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        stages = {k: float(np.mean([r.ms[k] for r in res])) for k in sequence.STAGES}
        line = {"metric": "end-to-end Mdisparities/s (stereo + triangulation + plane + xyzC in memory)",
                "value": a.frames * W * H / float(dt[0]) / 1e6, "unit": "Mdisp/s", "frames_per_s": a.frames / float(dt[0]),
                "n_gpus": world, "frames": a.frames, "mode": a.mode, "frames_in_flight": a.depth, "sms_per_sweep": a.sweep_workers, "ms_per_frame_per_gpu": float(dt[0]) * 1e3 / (a.frames / world),
                "stage_ms_host_clock": stages, "points_per_frame": int(np.mean([r.n_points for r in res])),
                "planes_valid": int(np.sum(~np.isnan(planes[:, 0]))), "mean_plane": [float(v) for v in mean]}
        print(json.dumps(line), flush=True)
    for x in h:
        x.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
