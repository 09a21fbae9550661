"""Frame-parallel launcher: one rank per GPU, frames (workdirs) round-robin over ranks, no traffic during
stereo, one all-reduce of the NaN-aware plane sums at the end (what wasscli + wassgridsurface do through
planes.txt and np.nanmean: cli/wasscli/wasscli.py:305-364, gridding/wassgridsurface/wassgridsurface.py:672-678).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m wass_b200.launcher --config config/stereo_config.txt --out output/planes.txt output/*_wd
"""
import argparse
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "bin", "wass_stereo")


def shard(n_items, rank, world):
    """Indices of the frames owned by `rank` (round-robin: frame i -> rank i % world)."""
    return list(range(rank, n_items, world))


def read_plane(workdir):
    """plane.txt -> 4 floats (nan when RANSAC failed or the frame crashed), wasscli.py:341-343."""
    try:
        with open(os.path.join(workdir, "plane.txt")) as f:
            v = [float(x) for x in f.read().split()]
        return np.array(v, np.float64) if len(v) == 4 else np.full(4, np.nan)
    except (OSError, ValueError):
        return np.full(4, np.nan)


def run_frame(config, workdir, device, exe=EXE):
    env = dict(os.environ, WASS_GPU_DEVICE=str(device))
    r = subprocess.run([exe, config, workdir], capture_output=True, text=True, env=env)
    return r.returncode, r.stdout


def plane_sums(planes):
    """NaN-aware accumulation through the C ABI (wsg_plane_mean_accumulate): [sum a, b, c, d, count]."""
    from . import capi
    return capi.plane_mean(planes)[1]


def reduce_planes(local_planes, n_frames, owned, dist=None):
    """All-reduce the plane sums and gather every frame's plane in frame order.
    Returns (mean_plane[4], planes[n_frames][4])."""
    import torch
    acc = torch.from_numpy(plane_sums(local_planes) if len(local_planes) else np.zeros(5))
    full = torch.full((n_frames, 4), float("nan"), dtype=torch.float64)
    for i, p in zip(owned, local_planes):
        full[i] = torch.from_numpy(np.asarray(p, np.float64))
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        acc = acc.to(dev)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        # every frame is owned by exactly one rank: nan elsewhere, so a NaN-ignoring max gathers them in order
        filled = torch.nan_to_num(full, nan=-float("inf")).to(dev)
        dist.all_reduce(filled, op=dist.ReduceOp.MAX)
        full = torch.where(torch.isinf(filled), torch.full_like(filled, float("nan")), filled).cpu()
        acc = acc.cpu()
    acc = acc.numpy()
    mean = acc[:4] / acc[4] if acc[4] > 0 else np.full(4, np.nan)
    return mean, full.numpy()


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--out", default=None, help="ordered planes.txt to write on rank 0")
    ap.add_argument("workdirs", nargs="+")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    owned = shard(len(a.workdirs), rank, world)
    planes = []
    for i in owned:
        rc, out = run_frame(a.config, a.workdirs[i], local)
        if rc != 0:
            print("[rank %d] %s failed (exit %d)\n%s" % (rank, a.workdirs[i], rc, out[-2000:]), flush=True)
        planes.append(read_plane(a.workdirs[i]))
    mean, allp = reduce_planes(planes, len(a.workdirs), owned, dist if world > 1 else None)
    if rank == 0:
        print("mean plane:", " ".join("%.17g" % v for v in mean))
        if a.out:
            with open(a.out, "w") as f:
                for p in allp:
                    f.write(" ".join("nan" if np.isnan(v) else "%.17g" % v for v in p) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
