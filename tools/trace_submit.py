import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, bench
from doom_rust_renderer_b200 import lib as drr, synth_wad
path, gm = bench.make_wad("e1m1")
views = bench.viewpoints(gm, "e1m1", 4096)
ctx = drr.Context(320, 200, 0, 4096); scene = drr.Scene(path, "E1M1", 320, 200); scene.upload_assets(ctx)
bench.record_batch(drr, ctx, scene, views, 3)
for i in range(3): ctx.submit(); ctx.sync()
os.environ["DRR_SUBMIT_TRACE"] = "1"
for ch in (1, 4, 8):
    os.environ["DRR_SUBMIT_CHUNKS"] = str(ch)
    t0 = time.perf_counter(); ctx.submit(); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    print("chunks", ch, "submit call %.3f ms, until sync %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), file=sys.stderr)
