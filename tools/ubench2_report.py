#!/usr/bin/env python3
"""Join the B200 log of tools/ubench2 (scheduler cycles per loop body) with the loop bodies' SASS (instruction counts per
opcode, read here with cuobjdump): issue rate of every body = instructions / cycles.  usage: ubench2_report.py ubench2 log"""
import re, subprocess, sys
exe, log = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True).stdout
bodies = {}
for part in re.split(r"\n\s*Function : ", sass)[1:]:
    name = re.match(r"_Z\d+(\w+?)Pj", part).group(1)
    ins = re.findall(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)[^;]*;", part)
    addr = {a: i for i, (a, _) in enumerate(ins)}
    # the loop: the last backward branch and its target
    best = None
    for i, (a, op) in enumerate(ins):
        m = re.search(r"/\*%s\*/[^;]*BRA[^;]* 0x([0-9a-f]+)" % a, part)
        if op.startswith("BRA") and m and m.group(1).zfill(4) in addr and addr[m.group(1).zfill(4)] < i:
            best = (addr[m.group(1).zfill(4)], i)
    if best:
        ops = {}
        for _, op in ins[best[0]:best[1] + 1]:
            k = op.split(".")[0]
            ops[k] = ops.get(k, 0) + 1
        bodies[name] = (best[1] - best[0] + 1, ops)
for line in open(log):
    m = re.match(r"(\w+)\s+([0-9.]+) scheduler cycles", line)
    if not m or m.group(1) not in bodies:
        continue
    n, ops = bodies[m.group(1)]
    cyc = float(m.group(2))
    print("%-22s %3d instr / %7.2f cycles = %.2f per cycle   %s" % (m.group(1), n, cyc, n / cyc, " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1]))))
