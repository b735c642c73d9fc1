#!/usr/bin/env python3
"""Per SOURCE LINE view of an .ncu-rep: joins ncu's per-instruction samples / executed counts (SASS page) with the line table
nvdisasm prints for the object file (built with -lineinfo).  usage: ncu_lines.py report.ncu-rep object.o kernel-substring [top]"""
import csv, io, re, subprocess, sys, tempfile, os, glob
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
kidx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(allrows) if r and r[0].startswith("Kernel Name")]
rows = allrows[starts[kidx]:(starts[kidx + 1] if kidx + 1 < len(starts) else len(allrows))]
kernel_full = rows[0][1]
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
ins = [r for r in rows[2:] if len(r) >= len(hdr)]
# the function's section in the disassembly: pick the one whose instruction count matches
best = None
for part in re.split(r"\n//-+ \.text\.", dis)[1:]:
    name = part.split(" ", 1)[0]
    if kname not in name:
        continue
    lines, cur = [], ("?", 0)
    for ln in part.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if best is None or abs(len(lines) - len(ins)) < abs(len(best) - len(ins)):
        best = lines
assert best is not None, "kernel not found in the object"
print(kernel_full[:100], "| %d instructions in the report, %d in the object" % (len(ins), len(best)))
agg = {}
tot_s = tot_i = 0
for r, ln in zip(ins, best):
    s, n = int(r[col["# Samples"]]), int(r[col["Instructions Executed"]])
    a = agg.setdefault(ln, [0, 0, {}])
    a[0] += s
    a[1] += n
    tot_s += s
    tot_i += n
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = int(r[col[k]])
            if v:
                a[2][k[6:]] = a[2].get(k[6:], 0) + v
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
srcs = {}
for ln, (s, n, why) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f = ln[0]
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "**", f), recursive=True)
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][ln[1] - 1].strip()[:90] if 0 < ln[1] <= len(srcs[f]) else ""
    reasons = " ".join("%s:%d%%" % (k, 100 * v // max(1, s)) for k, v in sorted(why.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %-90s %s" % (100.0 * s / tot_s, 100.0 * n / tot_i, f, ln[1], text, reasons))
