#!/usr/bin/env python
"""Per-opcode histogram of the HOT PATH of a kernel's inner loop, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_hist.py <object-or-library> <kernel-name-substring> [--loop N] [--list]

The inner loops of the sweep kernels contain bounded wait loops (progress-counter polls with nanosleep back-off) that are
skipped when the data is already there.  The hot path is therefore extracted by walking the loop body from its top and
TAKING every forward branch that jumps over a block containing NANOSLEEP (the slow path of a wait), falling through
everything else.  Loops are found as backward branches; --loop picks the N-th largest (default 0 = the largest body).
"""
import collections
import re
import subprocess
import sys


def sass_of(path, name):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    hits = [k for k in funcs if name in k]
    if not hits:
        sys.exit("no function matches %r; have e.g. %s" % (name, list(funcs)[:5]))
    return hits[0], funcs[hits[0]]


def opcode(text):
    t = re.sub(r"^@!?U?P\w+\s+", "", text)
    return t.split()[0]


def target(text):
    m = re.search(r"\b(?:BRA|BRA\.U|BRA\.DIV)\b.*?(0x[0-9a-f]+)\s*$", text)
    return int(m.group(1), 16) if m else None


def main():
    path, name = sys.argv[1], sys.argv[2]
    which = int(sys.argv[sys.argv.index("--loop") + 1]) if "--loop" in sys.argv else 0
    fname, ins = sass_of(path, name)
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if opcode(t).startswith("BRA") and not opcode(t).startswith("BRA.DIV"):
            tg = target(t)
            if tg is not None and tg < a and tg in addr:
                loops.append((i - addr[tg], addr[tg], i))
    loops.sort(reverse=True)
    if "--loops" in sys.argv:
        for k, (size, lo, hi) in enumerate(loops):
            body = [opcode(t) for _, t in ins[lo:hi + 1]]
            print("%2d  0x%05x..0x%05x  %5d instr  VIMNMX3 %3d  NANOSLEEP %2d" % (k, ins[lo][0], ins[hi][0], size + 1,
                  sum(o.startswith("VIMNMX3") for o in body), sum(o.startswith("NANOSLEEP") for o in body)))
        return
    # nested wait loops are small; the pixel loops are the big ones
    size, lo, hi = loops[which]
    hist = collections.Counter()
    path_ins = []
    i = lo
    while i <= hi:
        a, t = ins[i]
        op = opcode(t)
        tg = target(t) if op.startswith("BRA") and not op.startswith("BRA.DIV") else None
        hist[op] += 1
        path_ins.append((a, t))
        if tg is not None and tg > a and tg in addr and addr[tg] <= hi + 1:
            skipped = ins[i + 1:addr[tg]]
            if any("NANOSLEEP" in s for _, s in skipped):
                i = addr[tg]
                continue
        i += 1
    total = sum(hist.values())
    print("# %s" % fname)
    print("# loop 0x%x..0x%x: %d static instructions, %d on the hot path" % (ins[lo][0], ins[hi][0], size + 1, total))
    groups = collections.OrderedDict([
        ("packed int16 ALU (VIMNMX3/VIADDMNMX/VIMNMX/VIADD.16x2)", r"^(VIMNMX3|VIADDMNMX|VIMNMX|VIADD\.16)"),
        ("IMAD (FMA pipe: adds, moves, addressing)", r"^IMAD"),
        ("shuffle / warp reduce (SHFL, REDUX, CREDUX, VOTE)", r"^(SHFL|REDUX|CREDUX|VOTE|MATCH)"),
        ("select / permute / logic (SEL, PRMT, LOP3, SHF, POPC, LEA)", r"^(SEL|PRMT|LOP3|SHF|POPC|LEA|FLO|BREV)"),
        ("compare / integer add (ISETP, IADD3, VIADD, IABS)", r"^(ISETP|IADD|VIADD|IABS|I2F|F2I|FMUL|MUFU|FSETP|FADD|FFMA)"),
        ("shared memory (LDS, STS)", r"^(LDS|STS)"),
        ("global memory (LDG, STG, LD, ST, LDGSTS, ATOM, RED)", r"^(LDG|STG|LD\.|ST\.|LDGSTS|ATOM|RED|LDGDEPBAR|DEPBAR|UBLKCP|UTMA)"),
        ("control (BRA, BSSY, BSYNC, WARPSYNC, NOP, ...)", r"^(BRA|BSSY|BSYNC|WARPSYNC|NOP|YIELD|EXIT|BREAK|CALL|RET|NANOSLEEP|BAR)"),
    ])
    left = dict(hist)
    for gname, pat in groups.items():
        n = sum(c for op, c in hist.items() if re.match(pat, op))
        for op in [op for op in left if re.match(pat, op)]:
            del left[op]
        print("%5d  %s" % (n, gname))
    print("%5d  other: %s" % (sum(left.values()), ", ".join("%s %d" % kv for kv in sorted(left.items()))))
    print()
    for op, c in hist.most_common():
        print("%5d  %s" % (c, op))
    if "--list" in sys.argv:
        print()
        for a, t in path_ins:
            print("%06x  %s" % (a, t))


if __name__ == "__main__":
    main()
