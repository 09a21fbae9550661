#!/usr/bin/env python
"""bench.py -- headline benchmark of the wass_stereo dense-stereo hot path on B200.

Metric (BASELINE.json): Mdisparities/s (output disparity pixels W*H per second) on
2448x2048 pairs with 256 disparities, full 8-path SGM (cv::StereoSGBM MODE_HH arithmetic),
WASS default matcher parameters.  One "step" = one BATCH of rectified stereo pairs (--batch, default 16 per GPU) through the dense matcher
(prefilter -> cost volume -> 8-path aggregation -> WTA/LR/sub-pixel -> 3x3 median): one wsg_sgbm_compute_batch call.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by the driver through torch.distributed.run (one rank per GPU); frames shard one batch per
rank with no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os

# rank 0 prints ONE JSON line on stdout.  Libraries print there too (NCCL's "NCCL version ..." banner comes through C stdio
# whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to
# the real stdout at the end (emit()).
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG, NDISP = 2448, 2048, 256
WORKLOAD = "2448x2048 rectified pair, 256 disparities, MODE_HH 8-path, win 13, P1 338, P2 10816 (BASELINE configs[1])"


def wass_params(num_disp, mode):
    win = 13
    return dict(minDisparity=1, numDisparities=num_disp, blockSize=win, P1=2 * win * win, P2=64 * win * win,
                disp12MaxDiff=-1, preFilterCap=60, uniquenessRatio=1, speckleWindowSize=-70, speckleRange=16, mode=mode)


def ncu_traffic(frames_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json, written by
    tools/ncu_summary.py traffic), scaled from the capture's frames per launch to this run's: never measured under the
    timed run, so None when the file is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["sweep_kernel"]["dram_bytes_per_launch"]) * frames_per_launch / float(t.get("frames_per_launch", 1)), t.get("source")
    except Exception:
        return None, None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frame(seed):
    from wass_b200 import synth
    r, l, _ = synth.make_pair(W_IMG, H_IMG, NDISP, seed=seed)
    return synth.pad_for_sgbm(r, l, NDISP)


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic (cv2.StereoSGBM, the routine wass_stereo.cpp:837 calls)
# --------------------------------------------------------------------------------------------------
def _cpu_sgbm(img1, img2, p):
    try:
        import cv2
        cv2.setNumThreads(1)
        m = cv2.StereoSGBM_create(p["minDisparity"], p["numDisparities"], p["blockSize"], p["P1"], p["P2"])
        m.setUniquenessRatio(p["uniquenessRatio"]); m.setDisp12MaxDiff(p["disp12MaxDiff"])
        m.setPreFilterCap(p["preFilterCap"]); m.setSpeckleRange(p["speckleRange"])
        m.setSpeckleWindowSize(p["speckleWindowSize"])
        m.setMode(cv2.STEREO_SGBM_MODE_HH if p["mode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)
        return m.compute(img1, img2), "cv2.StereoSGBM %s" % cv2.__version__
    except ImportError:
        from oracle import sgbm
        return sgbm.compute(img1, img2, p)["disp"], "oracle/sgbm_oracle.c"


def _cpu_band_worker(args):
    seed, rows = args
    from wass_b200 import synth
    r, l, _ = synth.make_pair(W_IMG, rows, NDISP, seed=seed)
    i1, i2 = synth.pad_for_sgbm(r, l, NDISP)
    t = time.perf_counter()
    _cpu_sgbm(i1, i2, wass_params(NDISP, 1))
    return time.perf_counter() - t


def cpu_baseline_single(rows=2048):
    """1 core, `rows` rows of the benchmark frame (bounded sample: the whole frame, ~6 s on the GPU box's host)."""
    from wass_b200 import synth
    r, l, _ = synth.make_pair(W_IMG, rows, NDISP, seed=0)
    i1, i2 = synth.pad_for_sgbm(r, l, NDISP)
    t = time.perf_counter()
    disp, what = _cpu_sgbm(i1, i2, wass_params(NDISP, 1))
    dt = time.perf_counter() - t
    return {"value": W_IMG * rows / dt / 1e6, "unit": "Mdisp/s", "cores": 1, "kind": "port",
            "sample": "the whole %dx%d benchmark frame of seed 0 (D=256, MODE_HH) through %s, the routine "
                      "wass_stereo.cpp:837 calls; %.1f s; its disparity is the parity check of the GPU result" % (W_IMG, rows, what, dt)}, \
        (disp if rows == H_IMG else None)


def run_reference(args, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    P = os.cpu_count() or 1
    rows = 128
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_band_worker, [(s, rows) for s in range(P)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_band_worker, [(1000 * k + s, rows) for s in range(P)])
        dt = time.perf_counter() - t0
    value = P * args.steps * W_IMG * rows / dt / 1e6
    sample = "each step = %d bands of %dx%d (D=256, MODE_HH), one per host process, cv2.StereoSGBM 1 thread each" % (P, W_IMG, rows)
    line = {"impl": "reference", "metric": "Mdisparities/s", "value": value, "unit": "Mdisp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample, "frames_per_step_per_gpu": None, "handles_in_flight_per_gpu": None,
                       "rows_per_band": None,
                       "single_frame_ms": None, "padded_width": W_IMG + NDISP, "W1": W_IMG - 1, "l2": None,
                       "parallelism": "%d host processes" % P, "generator": "as the GPU arm (wass_b200/synth.py)",
                       "max_cost": None, "out_of_domain": None},
            "cpu_baseline": {"value": value, "unit": "Mdisp/s", "cores": P, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mdisp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def _sha(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def pinned_hashes():
    """tests/golden/fullsize_hashes.json: SHA-256 of the cv2.StereoSGBM disparity of the benchmark frames (seeds 0..2),
    made in the build container by tests/golden/make_fullsize_hashes.py."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize_hashes.json")) as f:
            j = json.load(f)
        return {c["seed"]: c for n, c in j["cases"].items() if n.startswith("bench_frame_seed")}, j.get("cv2")
    except Exception:
        return {}, None


def synthetic_plane(frame):
    """Stand-in per-frame sea plane for the plane-reduction leg of the multi-GPU run (the dense matcher alone fits no
    plane; tools/bench_sequence.py runs the whole frame pipeline with real fits).  Every 7th frame is a RANSAC failure."""
    if frame % 7 == 3:
        return [float("nan")] * 4
    th = 0.4 + 1e-3 * frame
    return [0.01 * np.sin(frame), np.sin(th), np.cos(th), -(8.0 + 0.01 * frame)]


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from wass_b200 import capi

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p = wass_params(NDISP, capi.MODE_HH)
    B = max(1, args.batch)
    # frames shard across GPUs: rank r owns seeds r*B .. r*B+B-1
    seeds = [rank * B + i for i in range(B)]
    frames = [make_frame(s) for s in seeds]
    H, Wp = frames[0][0].shape
    main = torch.cuda.Stream()
    torch.cuda.set_stream(main)
    i1 = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
    i2 = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    # ONE handle = one device arena (B cost + B aggregated volumes) + one compute stream (+ two copy streams for the
    # asynchronous host-buffer path)
    h = capi.Handle(local)
    if args.agg_impl >= 0:
        h.sgbm_set_impl(args.agg_impl)
    st = torch.cuda.Stream()
    h.set_stream(st.cuda_stream)
    dev_out = torch.empty((B, H, Wp), dtype=torch.int16, device="cuda")
    pin_in1 = [torch.from_numpy(f[0]).pin_memory() for f in frames]
    pin_in2 = [torch.from_numpy(f[1]).pin_memory() for f in frames]
    pin_out = [[torch.empty((H, Wp), dtype=torch.int16).pin_memory() for _ in frames] for _ in range(2)]      # one set per slot
    torch.cuda.synchronize()

    def step_dev():          # one step = one batch of B frames, device-resident inputs and outputs
        h.sgbm_compute_batch_device(B, i1.data_ptr(), i2.data_ptr(), H * Wp, H, Wp, Wp, p, dev_out.data_ptr())

    def submit(slot):        # the same through the host-buffer entry points: H2D + kernels + D2H, asynchronous
        h.sgbm_batch_submit(slot, B, [t.data_ptr() for t in pin_in1], [t.data_ptr() for t in pin_in2], H, Wp, Wp, p,
                            [t.data_ptr() for t in pin_out[slot]])

    def run_e2e(nsteps):
        """nsteps batches through wsg_sgbm_batch_submit / _wait, two in flight: the copies of one batch run beside the
        kernels of the other.  Returns when the last result is in host memory."""
        submit(0)
        for k in range(1, nsteps):
            submit(k & 1)
            h.sgbm_batch_wait((k - 1) & 1)
        h.sgbm_batch_wait((nsteps - 1) & 1)

    def timed_device(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        st.wait_event(e0)
        for _ in range(nsteps):
            step_dev()
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)          # clocks are sampled from the warm-up to the end of the e2e leg
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    barrier()
    # ---- per-stage profile: every kernel timed alone by CUDA events on the handle's stream (the roofline numbers)
    prof_steps = max(2, min(args.steps, 5))
    h.profile_enable(True)
    h.profile_reset()
    ms_serial = timed_device(prof_steps)
    prof = h.profile_get()
    h.profile_enable(False)
    stats = h.sgbm_stats()
    # one frame at a time, for the latency figure
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.sgbm_compute_batch_device(1, i1.data_ptr(), i2.data_ptr(), H * Wp, H, Wp, Wp, p, dev_out.data_ptr())
    h.synchronize()
    e0.record(st)
    for _ in range(3):
        h.sgbm_compute_batch_device(1, i1.data_ptr(), i2.data_ptr(), H * Wp, H, Wp, Wp, p, dev_out.data_ptr())
    e1.record(st)
    h.synchronize()
    ms_single = e0.elapsed_time(e1) / 3
    step_dev()

    # ---- device-resident throughput ("value"): K batches back to back
    ms_dev = timed_device(args.steps)

    # ---- end to end through the C ABI with HOST buffers (pinned)
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    # multi-GPU: the one collective of the path -- the NaN-aware mean of the per-frame planes -- inside the timed region,
    # through the C ABI (wsg_plane_allreduce: ncclAllReduce over NVLink).  The matcher alone fits no planes: synthetic ones.
    plane_info = None
    if world > 1:
        tp = time.perf_counter()
        mine = [synthetic_plane(rank * args.steps * B + j) for j in range(args.steps * B)]
        _, acc = capi.plane_mean(mine)
        mean, nfr = h.plane_allreduce(comm_holder["comm"], acc)
        plane_info = {"ms": (time.perf_counter() - tp) * 1e3, "mean": mean, "frames": nfr}
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- parity on the frames that were timed: device path == host path, and both == cv2
    outs = [pin_out[0][i].numpy() for i in range(B)]
    same = all(bool(torch.equal(dev_out[i].cpu(), pin_out[k][i])) for k in range(2) for i in range(B))

    if rank == 0:
        px = W_IMG * H_IMG
        value = world * args.steps * B * px / (ms_dev * 1e-3) / 1e6
        e2e = world * args.steps * B * px / (ms_e2e * 1e-3) / 1e6
        V = stats["volume_bytes"]
        rows_per_band = int(os.environ.get("WSG_SWEEP_ROWS", 0)) or "chosen per launch by the wave model of sweep_rows_per_band (sweep_kernels.cu)"
        nfr_prof = prof_steps * B
        agg_ms, agg_launches = prof["aggregate"]
        agg_ms_per_frame = agg_ms / nfr_prof
        peak, peak_src = measured_peak_gbs()
        alg_bytes = 4.0 * V                       # aggregation sweeps alone = 4*V per frame
        achieved = alg_bytes / (agg_ms_per_frame * 1e-3) / 1e9
        impl = stats["agg_impl"]
        if impl == 0:      # 8 single-direction launches (first 2V, others 3V); separate WTA reads V
            phys_bytes, kname, nlaunch = (2 + 3 * 7) * V, "aggregate_kernel (8 launches per frame)", 8.0 * B
        elif impl == 1:    # sweep 1: read C, write S; sweep 2: read C, read S, write S; separate WTA reads V
            phys_bytes, kname, nlaunch = 5 * V, "sweep_kernel (2 launches per batch), S written, separate WTA", 2.0
        else:              # sweep 1: read C, write S; sweep 2: read C, read S, WTA inside
            phys_bytes, kname, nlaunch = 4 * V, "sweep_kernel (2 launches per batch of %d frames), WTA fused into the second" % B, 2.0
        traffic, traffic_src = ncu_traffic(B) if impl >= 2 else (None, None)
        # parity block
        pins, cv2ver = pinned_hashes()
        checked, mism, how = 0, 0, []
        for i, s in enumerate(seeds):
            c = pins.get(s)
            if c and _sha(frames[i][0], frames[i][1]) == c["inputs_sha256"]:
                checked += outs[i].size
                mism += 0 if _sha(outs[i]) == c["disp_sha256"] else outs[i].size
                how.append("seed %d: sha256 of cv2 %s output (tests/golden/fullsize_hashes.json)" % (s, cv2ver))
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu, cpu_disp = cpu_baseline_single()
            if cpu_disp is not None:
                checked += cpu_disp.size
                mism += int((cpu_disp != outs[0]).sum())
                how.append("seed 0: every pixel against the cv2 run of cpu_baseline (same process, same frame)")
        parity = {"pixels": checked, "mismatches": mism, "against": "; ".join(how) if how else "nothing available",
                  "device_vs_host_path_bit_exact": same}
        line = {
            "metric": "Mdisparities/s", "value": value, "unit": "Mdisp/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": "each step = one batch of %d whole frames per GPU (seeds %d..%d on rank 0): wsg_sgbm_compute_batch_device for `value`, "
                                 "wsg_sgbm_batch_submit / _wait with pinned host buffers, two batches in flight, for `e2e`" % (B, seeds[0], seeds[-1]),
                       "frames_per_step_per_gpu": B, "handles_in_flight_per_gpu": 1, "rows_per_band": rows_per_band,
                       "single_frame_ms": ms_single, "padded_width": Wp, "W1": stats["width1"],
                       "l2": "inputs larger than L2 (C and S volumes %.2f GB each per frame, %d frames per batch)" % (V / 1e9, B),
                       "parallelism": "frame-per-GPU x%d" % world,
                       "generator": "grey range [80,176), noise sigma 2 (wass_b200/synth.py): softened from SURVEY 8d's full-range "
                                    "texture so that max(C)+P2 <= 32767, the domain in which cv2 is reproduced bit for bit",
                       "max_cost": stats["max_cost"], "out_of_domain": stats["out_of_domain"]},
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": kname, "agg_impl": impl, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "bound_note": "the sweeps are integer-ALU / issue bound, not HBM bound (DESIGN.md section 4); "
                         "frac is the HBM-roofline fraction BASELINE.json asks for",
                         "algorithmic_bytes_per_launch": alg_bytes * B / nlaunch,
                         "ms_per_launch": agg_ms_per_frame * B / nlaunch,
                         "algorithmic_bytes_per_frame": alg_bytes, "ms_per_frame": agg_ms_per_frame,
                         "frames_per_launch": B, "launches_per_batch": agg_launches / prof_steps,
                         "moved_bytes_per_frame_this_build": phys_bytes,
                         "timed": "CUDA events on the handle's stream around the sweep launches of %d batches run one at a time" % prof_steps},
            "stage_ms_per_frame": {k: v[0] / nfr_prof for k, v in prof.items() if v[1]},
            "batch_ms_one_at_a_time": ms_serial / prof_steps,
            "e2e": {"value": e2e, "unit": "Mdisp/s", "h2d_bytes_per_step": 2 * H * Wp * B, "d2h_bytes_per_step": 2 * H * Wp * B,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": stats["kernel_launches"] * args.steps,
            "clocks": clocks,
        }
        if plane_info is not None:
            allp = np.array([synthetic_plane(j) for j in range(world * args.steps * B)])
            ref = np.nanmean(allp, axis=0)
            line["plane_reduction"] = {"collective": "ncclAllReduce(sum, 5 x f64) through wsg_plane_allreduce, inside the e2e timed region",
                                       "frames": plane_info["frames"], "ms": plane_info["ms"],
                                       "max_abs_err_vs_numpy_nanmean": float(np.abs(plane_info["mean"] - ref).max()),
                                       "planes": "synthetic (the matcher alone fits none; tools/bench_sequence.py runs the whole frame pipeline)"}
            assert line["plane_reduction"]["max_abs_err_vs_numpy_nanmean"] < 1e-12 and plane_info["frames"] == int(np.isfinite(allp[:, 0]).sum())
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
        if mism or not same:
            raise SystemExit("PARITY FAILURE: %d of %d pixels differ from cv2 (device==host path: %s)" % (mism, checked, same))
    h.close()
    if world > 1:
        comm_holder["comm"].close()
        dist.destroy_process_group()


comm_holder = {}


def setup_comm(rank, world):
    """The library's own NCCL communicator (wsg_nccl_comm_create): rank 0's unique id travels over torch.distributed."""
    import torch
    import torch.distributed as dist
    from wass_b200 import capi
    local = int(os.environ.get("LOCAL_RANK", 0))
    uid = torch.zeros(capi.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    comm_holder["comm"] = capi.NcclComm(local, world, rank, bytes(uid.cpu().numpy().tobytes()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--batch", type=int, default=16, help="frames per step and GPU (one wsg_sgbm_compute_batch call); "
                    "device memory: batch x 5.7 GB")
    ap.add_argument("--agg-impl", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="-1 library default, 0 per-direction launches, 1 fused sweeps, 2 fused sweeps + fused WTA")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
            setup_comm(rank, world)
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
