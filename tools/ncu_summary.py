#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU): headline metrics + basic-block instruction shares of the top kernel.
usage: ncu_summary.py report.ncu-rep pixels_per_launch [min_share]"""
import csv, subprocess, sys, io
rep, pixels = sys.argv[1], float(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # which captured launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2 + kidx]))
print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "smsp__warps_active.avg.per_cycle_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in d: print("  %-70s %s %s" % (k, d[k], dict(zip(rows[0], rows[1])).get(k, "")))
for k in rows[0]:
    if "issue_stalled" in k and k.endswith("ratio") and float(d[k] or 0) > 0.2: print("  %-70s %s" % (k, d[k]))
ie = float(d["smsp__inst_executed.sum"])
print("  warp-instructions per pixel x32 (issue slots per pixel): %.2f" % (ie * 32 / pixels))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(allrows) if r and r[0].startswith("Kernel Name")]
lo = starts[kidx]
hi = starts[kidx + 1] if kidx + 1 < len(starts) else len(allrows)
rows = allrows[lo:hi]
hdr = rows[1]
ia, isrc, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
R = [(int(r[ia]), r[isrc].strip(), int(r[ith])) for r in rows[2:] if len(r) > ia]
tot = sum(n for n, _, _ in R)
start = 0
for i in range(1, len(R) + 1):
    if i == len(R) or R[i][0] != R[i - 1][0]:
        s = sum(n for n, _, _ in R[start:i])
        if s / tot > min_share:
            ops = {}
            for n, t, _ in R[start:i]:
                op = (t.split()[0] if not t.startswith("@") else t.split()[1]).split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            th = sum(t for _, _, t in R[start:i]) / max(1, s)
            print("[%d:%d] %5.1f%% slots/pix %6.2f n=%d cnt=%d thr=%.1f %s" % (start, i, s / tot * 100, s / pixels * 32, i - start, R[start][0], th,
                  " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1]))))
        start = i
