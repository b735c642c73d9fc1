"""Experiment: viewpoints -> checksums with TWO batches in flight (two contexts, two host threads) against one.
usage (GPU box): python tools/pipe_e2e.py [workload W H views phases]"""
import sys, time, threading, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import torch
import common
from common import drr, synth_wad
W, H, n, phases = 320, 200, 4096, 3
if len(sys.argv) > 1:
    W, H, n, phases = (int(a) for a in sys.argv[1:5])
path, gm = common.wad('e1m1')
views = np.array(synth_wad.walk_viewpoints(gm, n), np.float32)
def make():
    ctx = drr.Context(W, H, 0, n); scene = drr.Scene(path, 'E1M1', W, H); scene.upload_assets(ctx)
    st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
    return ctx, scene, st
C4 = [make() for _ in range(4)]; A, B = C4[0], C4[1]
def one(ctx, scene):
    ctx.reset(); scene.emit_views_device(ctx, views, 0.0, phases); ctx.draw(); return ctx.read_checksums(0, n)
ref = one(*A[:2]); assert (one(*B[:2]) == ref).all()
def loop(c, k, out):
    for _ in range(k): out.append(one(*c[:2]))
K = 64
for mode in ("one", "two", "one", "two"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if mode == "one":
        o = []; loop(A, K, o)
    else:
        oa, ob = [], []
        ta = threading.Thread(target=loop, args=(A, K // 2, oa)); tb = threading.Thread(target=loop, args=(B, K // 2, ob))
        ta.start(); tb.start(); ta.join(); tb.join(); o = oa + ob
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert all((x == ref).all() for x in o)
    print("%s in flight: %.4f ms per batch of %d views (%dx%d)" % (mode, dt / K * 1e3, n, W, H))
for k in (3, 4):
    for c in C4[:k]: assert (one(*c[:2]) == ref).all()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    outs = [[] for _ in range(k)]
    per = 96 // k
    ts = [threading.Thread(target=loop, args=(C4[i], per, outs[i])) for i in range(k)]
    [t.start() for t in ts]; [t.join() for t in ts]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%d in flight: %.4f ms per batch" % (k, dt / (per * k) * 1e3))
