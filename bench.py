#!/usr/bin/env python
"""bench.py -- headline benchmark of the wass_stereo dense-stereo hot path on B200.

Metric (BASELINE.json): Mdisparities/s (output disparity pixels W*H per second) on
2448x2048 pairs with 256 disparities, full 8-path SGM (cv::StereoSGBM MODE_HH arithmetic),
WASS default matcher parameters.  One "step" = one rectified stereo pair through the dense matcher
(prefilter -> cost volume -> 8-path aggregation -> WTA/LR/sub-pixel -> 3x3 median).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by the driver through torch.distributed.run (one rank per GPU); frames shard one per
rank with no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
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


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json,
    written by tools/ncu_summary.py traffic): never measured under the timed run, so None when the file is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["sweep_kernel"]["dram_bytes_per_launch"]), t.get("source")
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
    _, what = _cpu_sgbm(i1, i2, wass_params(NDISP, 1))
    dt = time.perf_counter() - t
    return {"value": W_IMG * rows / dt / 1e6, "unit": "Mdisp/s", "cores": 1, "kind": "port",
            "sample": "one %dx%d band (D=256, MODE_HH) of the benchmark frame through %s, the routine "
                      "wass_stereo.cpp:837 calls; %.1f s" % (W_IMG, rows, what, dt)}


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
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mdisp/s", "cores": P, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mdisp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from wass_b200 import capi

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p = wass_params(NDISP, capi.MODE_HH)
    depth = max(1, args.pipeline_depth)
    frames = [make_frame(depth * rank + i) for i in range(depth)]      # frames shard across GPUs, `depth` in flight per GPU
    H, Wp = frames[0][0].shape
    # One handle (= one device arena + one CUDA stream) per frame in flight.  The library launches on dedicated
    # (non-default) torch streams; torch events on a third stream bracket the timed region.
    main = torch.cuda.Stream()
    torch.cuda.set_stream(main)
    hs, streams, dev, pin = [], [], [], []
    for i in range(depth):
        h = capi.Handle(local)
        if args.agg_impl >= 0:
            h.sgbm_set_impl(args.agg_impl)
        st = torch.cuda.Stream()
        h.set_stream(st.cuda_stream)
        hs.append(h); streams.append(st)
        i1, i2 = frames[i]
        dev.append((torch.from_numpy(i1).cuda(), torch.from_numpy(i2).cuda(),
                    torch.empty((H, Wp), dtype=torch.int16, device="cuda")))
        pin.append((torch.from_numpy(i1).pin_memory(), torch.from_numpy(i2).pin_memory(),
                    torch.empty((H, Wp), dtype=torch.int16).pin_memory()))
    torch.cuda.synchronize()

    def step_dev(i):
        d1, d2, dd = dev[i]
        hs[i].sgbm_compute_device(d1.data_ptr(), d2.data_ptr(), H, Wp, Wp, p, dd.data_ptr())

    def step_e2e(i):
        p1, p2, pd = pin[i]
        hs[i].sgbm_compute_ptr(p1.data_ptr(), p2.data_ptr(), H, Wp, Wp, p, pd.data_ptr())

    def timed_device(nsteps, nstreams):
        """nsteps frames, frame k on handle k % nstreams; device time between two events on `main`."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        for st in streams[:nstreams]:
            st.wait_event(e0)
        for k in range(nsteps):
            step_dev(k % nstreams)
        for st in streams[:nstreams]:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    # ---- warm-up, then the per-stage profile of ONE frame at a time (kernels timed alone: roofline numbers)
    sampler = ClockSampler(local)          # clocks are sampled from the warm-up to the end of the e2e leg
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        for i in range(depth):
            step_dev(i)
    barrier()
    hs[0].profile_enable(True)
    hs[0].profile_reset()
    ms_serial = timed_device(args.steps, 1)
    prof = hs[0].profile_get()
    hs[0].profile_enable(False)
    stats = hs[0].sgbm_stats()

    # ---- device-resident throughput ("value"): `depth` frames in flight on as many streams.  With three or more in
    #      flight each sweep is confined to half the SMs (wsg_sgbm_set_sweep_workers): a wavefront over H/7 row bands
    #      leaves ~40 % of 148 workers waiting, so two frames' sweeps side by side move more pixels than one after the other.
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    workers = args.sweep_workers if args.sweep_workers >= 0 else (nsm // 2 if depth >= 3 else 0)
    for h in hs:
        h.sgbm_set_sweep_workers(workers)
    for i in range(depth):
        step_dev(i)
    ms_dev = timed_device(args.steps, depth)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernels + D2H per step, one host thread per
    #      frame in flight (the call is synchronous and releases the GIL)
    def e2e_worker(i, n):
        torch.cuda.set_device(local)
        for _ in range(n):
            step_e2e(i)

    def run_e2e(nsteps):
        ths = [threading.Thread(target=e2e_worker, args=(i, len(range(i, nsteps, depth)))) for i in range(depth)]
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()

    run_e2e(2 * depth)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    # cross-check of the two paths on this rank (bit-exact), every frame in flight
    same = all(bool(torch.equal(dev[i][2].cpu(), pin[i][2])) for i in range(depth))

    if rank == 0:
        px = W_IMG * H_IMG
        value = world * args.steps * px / (ms_dev * 1e-3) / 1e6
        e2e = world * args.steps * px / (ms_e2e * 1e-3) / 1e6
        V = stats["volume_bytes"]
        agg_ms, agg_launches = prof["aggregate"]
        agg_ms_per_frame = agg_ms / args.steps
        peak, peak_src = measured_peak_gbs()
        alg_bytes = 4.0 * V                       # aggregation sweeps alone = 4*V per frame
        achieved = alg_bytes / (agg_ms_per_frame * 1e-3) / 1e9
        impl = stats["agg_impl"]
        wta_ms = prof["wta"][0] / args.steps
        if impl == 0:      # 8 single-direction launches (first 2V, others 3V); separate WTA reads V
            phys_bytes, kname = (2 + 3 * 7) * V, "aggregate_kernel (8 launches/frame)"
        elif impl == 1:    # sweep 1: read C, write S; sweep 2: read C, read S, write S; separate WTA reads V
            phys_bytes, kname = 5 * V, "sweep_kernel (2 launches/frame), S written, separate WTA"
        elif impl in (2, 4):   # sweep 1: read C, write S; sweep 2: read C, read S, WTA inside
            phys_bytes, kname = 4 * V, "sweep_kernel (2 launches/frame), WTA fused into the second"
        else:              # 3-direction sweeps (2V, 2V) + two per-direction launches for the anti-diagonals (3V each)
            phys_bytes, kname = 10 * V, "sweep_kernel<NDIR=3> x2 + aggregate_kernel x2 (anti-diagonals), WTA fused into the last sweep"
        nlaunch = 2.0                             # sweep launches per frame (the reset kernel in the stage is ~13 us)
        traffic, traffic_src = ncu_traffic() if impl >= 2 else (None, None)
        line = {
            "metric": "Mdisparities/s", "value": value, "unit": "Mdisp/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "frames_in_flight_per_gpu": depth,
                       "sms_per_sweep_when_pipelined": workers if workers > 0 else nsm,
                       "single_frame_ms": ms_serial / args.steps, "padded_width": Wp, "W1": stats["width1"],
                       "l2": "inputs larger than L2 (C and S volumes %.2f GB each)" % (V / 1e9),
                       "parallelism": "frame-per-GPU x%d" % world, "device_vs_e2e_bit_exact": same,
                       "max_cost": stats["max_cost"], "out_of_domain": stats["out_of_domain"]},
            "roofline": {"bound": "hbm", "kernel": kname, "agg_impl": impl, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "bound_note": "the sweeps are integer-ALU bound, not HBM bound (DESIGN.md section 4); "
                         "frac is the HBM-roofline fraction BASELINE.json asks for",
                         "algorithmic_bytes_per_launch": alg_bytes / nlaunch if impl != 0 else alg_bytes / 8,
                         "ms_per_launch": agg_ms_per_frame / (nlaunch if impl != 0 else 8),
                         "algorithmic_bytes_per_frame": alg_bytes, "ms_per_frame": agg_ms_per_frame,
                         "launches_per_frame": agg_launches / args.steps,
                         "moved_bytes_per_frame_this_build": phys_bytes,
                         "moved_gbs": phys_bytes / (agg_ms_per_frame * 1e-3) / 1e9},
            "stage_ms_per_frame": {k: v[0] / args.steps for k, v in prof.items() if v[1]},
            "e2e": {"value": e2e, "unit": "Mdisp/s", "h2d_bytes_per_step": 2 * H * Wp, "d2h_bytes_per_step": 2 * H * Wp,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": stats["kernel_launches"] * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    for h in hs:
        h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--pipeline-depth", type=int, default=3,
                    help="frames in flight per GPU (one handle + stream each); 1 = one frame at a time")
    ap.add_argument("--sweep-workers", type=int, default=-1,
                    help="SMs per fused sweep while frames are pipelined (0 = all; -1 = half the SMs when 3+ frames are in flight)")
    ap.add_argument("--agg-impl", type=int, default=-1, choices=[-1, 0, 1, 2, 3, 4],
                    help="-1 library default, 0 per-direction launches, 1 fused sweeps, 2 fused sweeps + fused WTA, "
                         "3 three-direction sweeps + anti-diagonal launches + fused WTA")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
