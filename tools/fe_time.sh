#!/bin/bash
# front-end kernel time and viewpoints -> checksums time per pass (run under gpurun): tools/fe_time.sh "walk320 things640" [label]
for wl in $1; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
  python - "$wl" "$2" <<'P'
import json, sys
d = json.load(open("/tmp/s.json")); e = d["e2e"]; r = d["roofline"]
print("%-10s %-24s fe %.4f compact %.4f bin %.4f tile %.4f e2e %.4f ms/pass  parity %s" % (sys.argv[1], sys.argv[2], e["front_end_kernel_ms"], e["compaction_or_count_ms"], r["bin_kernel_ms"], r["kernel_ms"], e["ms_per_pass"], d.get("parity_ok")))
P
done
