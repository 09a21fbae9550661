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

_REAL_STDOUT = os.dup(1)     # keep stdout to the one JSON line: libraries (NCCL banner) print to fd 1 too
os.dup2(2, 1)
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
    ap.add_argument("--depth", type=int, default=2, help="handles per GPU (one arena, stream and host thread each)")
    ap.add_argument("--batch", type=int, default=8, help="frames per batched matcher run (wsg_dense_stereo_batch)")
    ap.add_argument("--c-abi-nccl", type=int, default=1, help="1: reduce the planes with wsg_plane_allreduce (the library's own NCCL "
                    "communicator), 0: torch.distributed")
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
    xyzc_out = [[torch.empty(148 + 6 * W * H, dtype=torch.uint8).pin_memory().numpy() for _ in range(a.batch)]
                for _ in range(a.depth)]   # reusable pinned destinations
    comm = None
    if world > 1 and a.c_abi_nccl:
        uid = torch.zeros(capi.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        comm = capi.NcclComm(local, world, rank, uid.cpu().numpy().tobytes())
    sequence.run_sequence(frames[: a.batch * a.depth * world], calib, dense, device=local, rank=rank, world=world, dist=None, handle=h,
                          xyzc_out=xyzc_out, batch=a.batch)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean, planes, res = sequence.run_sequence(frames, calib, dense, device=local, rank=rank, world=world,
                                              dist=dist if world > 1 and comm is None else None, handle=h, xyzc_out=xyzc_out,
                                              batch=a.batch)
    t_red = 0.0
    if comm is not None:
        # the final plane reduction through the C ABI: ncclAllReduce(sum) over 5 doubles + an all-gather of the per-frame
        # planes for an ordered planes.txt, both inside the timed region
        tr = time.perf_counter()
        mine = np.array([r.plane for r in res])
        mean, nvalid = h[0].plane_allreduce(comm, capi.plane_mean(mine)[1])
        per = -(-a.frames // world)
        padded = np.full((per, 4), np.nan)
        padded[:mine.shape[0]] = mine
        allp = h[0].plane_allgather(comm, padded)                 # [rank][j] = frame j*world + rank
        planes = np.full((a.frames, 4), np.nan)
        for r in range(world):
            idx = np.arange(r, a.frames, world)
            planes[idx] = allp[r, :idx.size]
        t_red = (time.perf_counter() - tr) * 1e3
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        stages = {k: float(np.mean([r.ms[k] for r in res])) for k in sequence.STAGES}
        line = {"metric": "end-to-end Mdisparities/s (stereo + triangulation + plane + xyzC in memory)",
                "value": a.frames * W * H / float(dt[0]) / 1e6, "unit": "Mdisp/s", "frames_per_s": a.frames / float(dt[0]),
                "n_gpus": world, "frames": a.frames, "mode": a.mode, "handles_per_gpu": a.depth, "frames_per_batch": a.batch,
                "plane_reduction": ("wsg_plane_allreduce + wsg_plane_allgather (NCCL through the C ABI), %.2f ms" % t_red) if comm is not None else ("torch.distributed" if world > 1 else "local"),
                "ms_per_frame_per_gpu": float(dt[0]) * 1e3 / (a.frames / world),
                "stage_ms_host_clock": stages, "points_per_frame": int(np.mean([r.n_points for r in res])),
                "planes_valid": int(np.sum(~np.isnan(planes[:, 0]))), "mean_plane": [float(v) for v in mean]}
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    for x in h:
        x.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
