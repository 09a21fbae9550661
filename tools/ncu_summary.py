#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
    python tools/ncu_summary.py launches <launches.csv> <out.txt>
    python tools/ncu_summary.py rep <file.ncu-rep> <out.txt>
    python tools/ncu_summary.py traffic <file.ncu-rep> <out.json> "<note>" [frames per launch]
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
           "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write("%-72s %6s %12s %8s\n" % ("kernel", "n", "avg_us", "share%"))
        for n, v in agg.items():
            f.write("%-72s %6d %12.1f %8.1f\n" % (n, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(out).read())


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none: %s\n" % path)
        for r in rows[2:]:
            f.write("\n== %s\n" % r[H.index("Kernel Name")])
            for m in METRICS:
                if m in H:
                    f.write("%-80s %s %s\n" % (m, r[H.index(m)], units[H.index(m)]))
    print(open(out).read())


def traffic(rep, out_json, source_note, frames_per_launch=1):
    """profiles/traffic.json: DRAM bytes per launch (read+write, averaged over the captured launches) per kernel family."""
    import json
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H = rows[0]
    ki, ri, wi, ti = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
    units = rows[1]
    def scale(u):
        return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1.0)
    fam = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("wsg::", "").strip()
        rd = float(r[ri].replace(",", "")) * scale(units[ri]); wr = float(r[wi].replace(",", "")) * scale(units[wi])
        t = float(r[ti].replace(",", "")) * scale(units[ti])
        fam.setdefault(name, []).append((rd, wr, t))
    out = {"source": source_note, "frames_per_launch": frames_per_launch}
    for n, v in fam.items():
        out[n] = {"launches_captured": len(v), "dram_read_bytes_per_launch": sum(x[0] for x in v) / len(v),
                  "dram_write_bytes_per_launch": sum(x[1] for x in v) / len(v),
                  "dram_bytes_per_launch": sum(x[0] + x[1] for x in v) / len(v),
                  "ncu_duration_ms_per_launch": 1e3 * sum(x[2] for x in v) / len(v)}
    json.dump(out, open(out_json, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "", int(sys.argv[5]) if len(sys.argv) > 5 else 1)
    else:
        {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
