#!/usr/bin/env python3
"""Debug helper: render walk viewpoints with the current kernel, compare with the oracle, describe differing pixels."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import common
from common import drr, orc, synth_wad

W, H, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
path, gm = common.wad("e1m1")
game = orc.Game(path, "E1M1", W, H)
views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[:: max(1, 4096 // (n + 4))], n)
ctx = drr.Context(W, H, 0, n)
scene = drr.Scene(path, "E1M1", W, H)
scene.upload_assets(ctx)
scene.emit_views(ctx, views)
ctx.submit(); ctx.sync()
spans = ctx._list(3, drr.SPAN_DTYPE); colidx = ctx._list(4, drr.COLIDX_DTYPE); segs = ctx._list(1, drr.SEG_DTYPE); planes = ctx._list(2, drr.PLANE_DTYPE)
for k, v in enumerate(views):
    ref = game.render(float(v[0]), float(v[1]), float(v[2]))
    got = ctx.read_framebuffer(k)
    ys, xs = np.nonzero((got != ref).any(2))
    for y, x in list(zip(ys, xs))[:6]:
        ci = colidx[k * W + x]
        print("view", k, "x", x, "y", y, "got", got[y, x], "want", ref[y, x])
        for j in range(ci["n_opaque"] + ci["n_masked"]):
            s = spans[ci["first"] + j]
            if s["y0"] <= y <= s["y1"]:
                tag = "opaque" if j < ci["n_opaque"] else "masked"
                print("   ", tag, dict(kind=int(s["kind"]), y0=int(s["y0"]), y1=int(s["y1"]), top_y=int(s["top_y"]), bottom_y=int(s["bottom_y"])),
                      segs[s["op"]] if s["kind"] < 2 else planes[s["op"]])
