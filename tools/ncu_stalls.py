#!/usr/bin/env python3
"""Where the warps of the top kernel wait: stall samples of an .ncu-rep (source page) summed per basic block, with the
heaviest instructions of each.  usage: ncu_stalls.py report.ncu-rep [top_blocks] [launch index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(allrows) if r and r[0].startswith("Kernel Name")]
rows = allrows[starts[kidx]:(starts[kidx + 1] if kidx + 1 < len(starts) else len(allrows))]
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
R = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    R.append({"src": r[col["Source"]].strip(), "n": int(r[col["Instructions Executed"]]), "samples": int(r[col["# Samples"]]),
              "stalls": {s: int(r[col[s]]) for s in stall_cols}, "conf": r[col["L1 Wavefronts Shared Excessive"]]})
tot = sum(x["samples"] for x in R)
blocks, start = [], 0
for i in range(1, len(R) + 1):
    if i == len(R) or R[i]["n"] != R[i - 1]["n"]:
        blocks.append((start, i, sum(x["samples"] for x in R[start:i])))
        start = i
print("total samples", tot)
for lo, hi, s in sorted(blocks, key=lambda b: -b[2])[:top]:
    agg = {}
    for x in R[lo:hi]:
        for k, v in x["stalls"].items():
            agg[k] = agg.get(k, 0) + v
    reasons = " ".join("%s:%.0f%%" % (k[6:], 100.0 * v / max(1, s)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:5])
    print("[%d:%d] %.1f%% of samples, %d instr, cnt %d: %s" % (lo, hi, 100.0 * s / tot, hi - lo, R[lo]["n"], reasons))
    for x in sorted(R[lo:hi], key=lambda x: -x["samples"])[:6]:
        why = max(x["stalls"].items(), key=lambda kv: kv[1])
        print("      %5.2f%%  %-60s %s" % (100.0 * x["samples"] / tot, x["src"][:60], why[0][6:]))
