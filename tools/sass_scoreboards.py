#!/usr/bin/env python
"""Scoreboard use of a kernel's long-latency loads, decoded from the control bits of `cuobjdump -sass` (no GPU needed).

    python tools/sass_scoreboards.py <object-or-library> <kernel-name-substring> [opcode-regex, default LDG]

sm_70+ control word (bits 105..125 of the 128-bit instruction): stall count (4), yield (1), write barrier (3: the
scoreboard the result signals, 7 = none), read barrier (3), wait mask (6: scoreboards the instruction waits for).
A scoreboard is a counter and an instruction can only wait for it to reach ZERO: when two outstanding loads share one,
the consumer of the older load also waits for the younger one -- its full memory latency, however far ahead it was issued.
Lists every matching load with its scoreboard, and for every scoreboard the instructions that wait on it.
"""
import re
import subprocess
import sys


def main():
    path, name = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"^LDG")
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, ins = None, {}
    lines = out.splitlines()
    i = 0
    while i < len(lines):
        m = re.search(r"Function : (\S+)", lines[i])
        if m:
            cur = m.group(1); ins[cur] = []
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and cur and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                ins[cur].append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=ctrl & 15, yld=(ctrl >> 4) & 1,
                                     wr=(ctrl >> 5) & 7, rd=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 63))
                i += 1
        i += 1
    hits = [k for k in ins if name in k]
    if not hits:
        sys.exit("no function matches; have e.g. %s" % list(ins)[:4])
    f = ins[hits[0]]
    print("#", hits[0], len(f), "instructions")
    op = lambda t: re.sub(r"^@!?U?P\w+\s+", "", t)
    loads = [x for x in f if pat.match(op(x["text"]))]
    for x in loads:
        print("%06x  SB%d  %s" % (x["addr"], x["wr"], x["text"][:80]))
    print()
    for sb in sorted(set(x["wr"] for x in loads if x["wr"] != 7)):
        users = [x for x in f if x["wr"] == sb]
        kinds = {}
        for x in users:
            k = op(x["text"]).split()[0]
            kinds[k] = kinds.get(k, 0) + 1
        waits = sum(1 for x in f if x["wait"] >> sb & 1)
        print("SB%d: signalled by %s; %d instructions wait on it" % (sb, ", ".join("%s x%d" % kv for kv in sorted(kinds.items())), waits))


if __name__ == "__main__":
    main()
