#!/usr/bin/env python
"""Timeline of the bands of the LAST sweep of a batch, from the per-ticket timestamps the sweep kernel records under
WSG_SWEEP_DEBUG=1 (SM id, start, end of every band; printed by wsg_sgbm_get_stats / wsg_check_sweep to stderr).

    WSG_SWEEP_DEBUG=1 python tools/band_timeline.py [--n 8] [--size 2448x2048x256]

Prints: launch span, mean band duration, the busy fraction of the workers (sum of band durations / (workers x span)),
how late the first-wave bands start (the fill of the wavefronts) and how early the workers run dry (the drain)."""
import argparse
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--size", default="2448x2048x256")
    a = ap.parse_args()
    env = dict(os.environ, WSG_SWEEP_DEBUG="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep_probe.py"), "--n", str(a.n), "--reps", "1", "--size", a.size],
                       capture_output=True, text=True, env=env)
    rows = [l.split()[1:] for l in r.stderr.splitlines() if re.match(r"\[wsg\] \d+ \d+ \d+ \d+ ", l)]
    n_t = max(int(x[0]) for x in rows) + 1
    t = np.array([[float(v) for v in x] for x in rows[-n_t:]])       # the last dump = the timed run's last sweep
    ticket, frame, band, sm, t0, t1 = t.T
    span = t1.max() - t0.min()
    dur = t1 - t0
    workers = len(set(sm.astype(int)))
    first = np.argsort(ticket)[:workers]
    out = {"n": a.n, "tickets": int(n_t), "workers": workers, "span_us": round(float(span), 1), "band_us_mean": round(float(dur.mean()), 1),
           "band_us_p10_p90": [round(float(np.percentile(dur, 10)), 1), round(float(np.percentile(dur, 90)), 1)],
           "busy_fraction": round(float(dur.sum() / (workers * span)), 3),
           "first_wave_start_us_mean_max": [round(float((t0[first] - t0.min()).mean()), 1), round(float((t0[first] - t0.min()).max()), 1)],
           "last_end_minus_worker_end_us_mean": round(float(np.mean([t1.max() - t1[sm == s].max() for s in set(sm)])), 1)}
    # time a band spends before its first row can start is inside its duration: compare the first wave's durations
    out["band_us_first_wave_mean"] = round(float(dur[first].mean()), 1)
    out["band_us_later_mean"] = round(float(np.delete(dur, first).mean()), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
