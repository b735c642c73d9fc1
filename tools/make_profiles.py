#!/usr/bin/env python3
"""Turn the captures of tools/profile_final_r2b.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.
usage: make_profiles.py <tag in gpurun_out> <prefix in profiles>     e.g. make_profiles.py r2b_final r2b_final"""
import csv, json, os, shutil, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"

def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout

def summary(rep, pixels, dst):
    if not os.path.exists(rep):
        print("missing", rep); return
    txt = run(["python", "tools/ncu_summary.py", rep, str(pixels), "0.02"]) + run(["python", "tools/ncu_stalls.py", rep, "8"])
    open(dst, "w").write("\n".join(l[:260] for l in txt.splitlines()) + "\n")
    print("wrote", dst)

summary("%s/%s_prof_1280.ncu-rep" % (G, tag), 128 * 1280 * 800, "%s/%s_tile_1280x800.txt" % (P, out))
summary("%s/%s_prof_320.ncu-rep" % (G, tag), 1024 * 320 * 200, "%s/%s_tile_320x200.txt" % (P, out))
summary("%s/%s_prof_bin_320.ncu-rep" % (G, tag), 1024 * 320 * 200, "%s/%s_bin_320x200.txt" % (P, out))
if os.path.exists("%s/%s_launches.csv" % (G, tag)):
    shutil.copy("%s/%s_launches.csv" % (G, tag), "%s/%s_launches.csv" % (P, out))
fe = "%s/%s_fe_prof.ncu-rep" % (G, tag)
if os.path.exists(fe):
    txt = run(["python", "tools/ncu_lines.py", fe, "doom_rust_renderer_b200/csrc/build/drr_frontend.o", "frontend_kernelILb1ELi4ELi4", "45"])
    raw = list(csv.reader(run(["ncu", "-i", fe, "--page", "raw", "--csv"]).splitlines()))
    d = dict(zip(raw[0], raw[2]))
    head = ["warm caches (--cache-control none), walk320, 4096 viewpoints"]
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
              "smsp__warps_active.avg.per_cycle_active", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
              "lts__t_sector_hit_rate.pct"):
        head.append("  %-70s %s" % (k, d.get(k)))
    open("%s/%s_frontend_lines.txt" % (P, out), "w").write("\n".join(head) + "\n" + "\n".join(l[:230] for l in txt.splitlines()) + "\n")
    print("wrote frontend lines")
# DRAM traffic of one tile-kernel launch over the full batch
traffic = {}
views = {"walk320": 4096, "walk1280": 512, "walls1280": 256, "flats1280": 256, "things640": 4096, "stress1920": 1024}
for wl, n in views.items():
    path = "%s/%s_traffic_%s.csv" % (G, tag, wl)
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    m = {r[12]: float(r[14].replace(",", "")) for r in rows}
    traffic[wl] = {"views": n, "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
                   "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], "kernel": rows[0][4], "grid": rows[0][8],
                   "duration_under_ncu_ns": m["gpu__time_duration.sum"]}
if traffic:
    json.dump(traffic, open("%s/r2_traffic.json" % P, "w"), indent=1)
    print("wrote r2_traffic.json", {k: round(v["dram_bytes_per_launch"] / 1e6) for k, v in traffic.items()})
