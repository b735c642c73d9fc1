for v in 256 1024 4096 8192 16384; do
  python bench.py --workload walk320 --views $v --steps 5 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
  python - $v <<'P'
import json, sys
d = json.load(open("/tmp/s.json")); e = d["e2e"]
print("views %6s fe %.4f ms  e2e %.4f ms/pass" % (sys.argv[1], e["front_end_kernel_ms"], e["ms_per_pass"]))
P
done
