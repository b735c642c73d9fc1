#!/usr/bin/env python3
"""Render viewpoints of a WAD map on the GPU through the C ABI and write them as PNG files (needs a B200; run under gpurun).

    python tools/render_png.py [--wad PATH --map E1M1] [--size 1280x800] [--views N] [--out gpurun_out/frames]
                               [--tic T --seed S] [--host-front-end]

--tic puts the world T game ticks (1/35 s) after the start (light effects, sprite animation; synthetic WAD: the e1m1_time
variant, which has light-effect sectors).  The renderer's front-end runs on the GPU unless --host-front-end is given.

Without --wad the deterministic synthetic E1M1-class IWAD is used (no WAD file ships with the image)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from doom_rust_renderer_b200 import lib as drr, png, synth_wad  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--wad")
ap.add_argument("--map", default="E1M1")
ap.add_argument("--size", default="640x400")
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "frames"))
ap.add_argument("--tic", type=int, default=0)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--host-front-end", action="store_true")
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
os.makedirs(args.out, exist_ok=True)
if args.wad:
    path = args.wad
    scene = drr.Scene(path, args.map, W, H)
    x, y, a = scene.player_start()
    views = np.array([(x, y, a + 0.4 * k) for k in range(args.views)], np.float32)
else:
    data, gm, _ = synth_wad.build_wad("e1m1_time" if args.tic else "e1m1")
    path = os.path.join(args.out, "synth_e1m1.wad")
    open(path, "wb").write(data)
    scene = drr.Scene(path, "E1M1", W, H)
    walk = synth_wad.walk_viewpoints(gm, 4096)
    views = np.concatenate([np.array([scene.player_start()], np.float32), walk[:: max(1, 4096 // max(1, args.views - 1))][: args.views - 1]])
ctx = drr.Context(W, H, 0, len(views))
scene.upload_assets(ctx)
ts = args.tic / 35.0
if args.tic:
    scene.set_tic(args.tic, args.seed)
if args.host_front_end:
    skipped = scene.emit_views(ctx, views, timestamp=ts)
    ctx.submit()
else:
    skipped = scene.emit_views_device(ctx, views, timestamp=ts)
    ctx.draw()
ctx.sync()
crcs = ctx.read_checksums(0, len(views))
for k in range(len(views)):
    if k in skipped:
        continue
    name = os.path.join(args.out, "frame_%03d_%dx%d.png" % (k, W, H))
    png.write_png(name, ctx.read_framebuffer(k))
    print(name, "checksum %016x" % int(crcs[k]))
