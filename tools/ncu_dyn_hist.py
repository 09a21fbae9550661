#!/usr/bin/env python
"""Dynamic (executed) per-opcode histogram of a kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_dyn_hist.py <file.ncu-rep> <kernel-name-substring> [pixels]

Sums `Instructions Executed` (warp-level) per opcode over the kernel's SASS; with `pixels` given, prints the
count per pixel-row step (one warp = one pixel): the figure the pipe-floor model in DESIGN.md is built from.
Also lists the hottest instructions by stall samples.
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, name = sys.argv[1], sys.argv[2]
    pixels = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kernels, cur, hdr = {}, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = r[1]; kernels[cur] = []; hdr = None
        elif r and r[0] == "Address":
            hdr = r
        elif cur and hdr and len(r) == len(hdr):
            kernels[cur].append(dict(zip(hdr, r)))
    hits = [k for k in kernels if name in k.replace("(int)", "").replace("(bool)", "").replace(" ", "")]
    if not hits:
        sys.exit("no kernel matches; have: %s" % list(kernels))
    k = hits[0]
    ins = kernels[k]
    hist, stall = collections.Counter(), collections.Counter()
    total = 0
    for d in ins:
        op = re.sub(r"^@!?U?P\w+\s+", "", d["Source"].strip()).split()[0]
        n = int(d["Instructions Executed"])
        hist[op] += n; total += n
        stall[(d["Address"][-5:], d["Source"].strip()[:70])] += int(d["Warp Stall Sampling (All Samples)"] or 0)
    print("# %s" % k)
    print("# %d warp-instructions executed%s" % (total, ", %.1f per pixel-row step" % (total / pixels) if pixels else ""))
    classes = [("ALU half-rate (VIMNMX3, VIADDMNMX, PRMT, SEL, LOP3, SHF, ISETP, VIADD, IADD3, LEA, POPC)",
                r"^(VIMNMX3|VIADDMNMX|PRMT|SEL|LOP3|SHF|ISETP|VIADD|IADD3|LEA|POPC|IABS|PLOP3|FLO|VIADD)"),
               ("ALU 2-input min/max (VIMNMX)", r"^VIMNMX(?!3)"),
               ("FMA pipe (IMAD*, FFMA, FMUL)", r"^(IMAD|FFMA|FMUL|FADD)"),
               ("shuffle / reduce / vote", r"^(SHFL|REDUX|CREDUX|VOTE|MATCH)"),
               ("shared memory", r"^(LDS|STS)"),
               ("global memory / async copies", r"^(LDG|STG|LD\.|ST\.|LDGSTS|ATOM|RED|LDGDEPBAR|DEPBAR|UBLKCP|UTMA|LD$|ST$)"),
               ("branches / convergence", r"^(BRA|BSSY|BSYNC|WARPSYNC|NOP|YIELD|EXIT|BREAK|CALL|RET|NANOSLEEP|BAR|ENDCOLL)"),
               ("uniform datapath / constants / conversions", r"^(U[A-Z]|LDC|S2R|S2UR|CS2R|MOV|R2UR|I2F|F2I|MUFU)")]
    left = dict(hist)
    for cname, pat in classes:
        n = 0
        for op in list(left):
            if re.match(pat, op):
                n += left.pop(op)
        print("%14d %6.1f%% %s  %s" % (n, 100.0 * n / total, ("%7.1f/px" % (n / pixels)) if pixels else "", cname))
    print("%14d %6.1f%% %s  other: %s" % (sum(left.values()), 100.0 * sum(left.values()) / total, "", sorted(left)))
    print()
    for op, n in hist.most_common(40):
        print("%14d %6.1f%% %s  %s" % (n, 100.0 * n / total, ("%7.1f/px" % (n / pixels)) if pixels else "", op))
    print("\n# hottest instructions by stall samples")
    tot = sum(stall.values())
    for (a, s), n in stall.most_common(25):
        print("%6.2f%%  %s  %s" % (100.0 * n / max(tot, 1), a, s))


if __name__ == "__main__":
    main()
