#!/usr/bin/env python
"""List the loops (backward branches) of a SASS dump with their instruction mix.
usage: cuobjdump -sass -fun <kernel> lib.so | python tools/sass_loops.py [min_instructions]"""
import re, sys, collections
min_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if not m or "BRA.DIV" in t:
        continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr_idx:
        body = ins[addr_idx[tgt]:i + 1]
        if len(body) < min_n:
            continue
        c = collections.Counter()
        for _, s in body:
            op = s.split()[1] if s.startswith("@") else s.split()[0]
            c[op.split(".")[0]] += 1
        fp64 = sum(v for k, v in c.items() if k in FP64)
        print("loop 0x%05x..0x%05x  %4d instr  fp64-pipe %3d  | " % (tgt, a, len(body), fp64) +
              ", ".join("%s %d" % kv for kv in c.most_common(14)))
