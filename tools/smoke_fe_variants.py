"""Small device front-end + draw runs through every route of drr_fe_emit_views (single pass, forced two-pass, slab overflow,
enlarged working arrays), all phases, both maps -- a quick GPU smoke (`python tools/smoke_fe_variants.py` under gpurun).
(compute-sanitizer is closed on this GPU pool; bounds are argued in DESIGN.md section 8 and checked by the capacity guards.)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import common
from common import drr, synth_wad
for kind, W, H, n, env in (("e1m1", 320, 200, 96, {}), ("stress", 200, 120, 64, {}), ("e1m1", 324, 200, 40, {"DRR_FE_TWO_PASS": "1"}),
                           ("stress", 640, 400, 24, {"DRR_FE_CAP_RENDERS": "64", "DRR_FE_CAP_DSEGS": "8"})):
    for k in ("DRR_FE_TWO_PASS", "DRR_FE_CAP_RENDERS", "DRR_FE_CAP_DSEGS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    path, gm = common.wad(kind)
    views = np.array(synth_wad.walk_viewpoints(gm, n) if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, n), np.float32)
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    skipped = scene.emit_views_device(ctx, views, 0.0, 7)
    ctx.draw()
    crc = ctx.read_checksums(0, n)
    print(kind, W, H, n, env, "mode", ctx.fe_last_mode(), "skipped", len(skipped), "distinct frames", len(set(crc.tolist())))
