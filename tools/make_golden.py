#!/usr/bin/env python3
"""Freeze the oracle's output on the deterministic synthetic WAD into tests/golden/oracle_frames.json.

The reference has no golden vectors of its own and cannot be run here, so these vectors pin the
ORACLE (and therefore detect any later drift of the oracle, of the WAD generator or of libm), not the reference."""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from common import drr, orc, synth_wad  # noqa: E402

out = {"wad_crc32": {}, "frames": []}
for kind in ("e1m1", "stress", "tiny"):
    data, _, _ = synth_wad.build_wad(kind)
    out["wad_crc32"][kind] = zlib.crc32(data)
path, gm = common.wad("e1m1")
views = synth_wad.walk_viewpoints(gm, 4096)
plan = [(320, 200, ["spawn", 0, 511, 1023, 2047, 3000, 4095]), (640, 400, ["spawn", 700, 2500]), (1280, 800, [1234]), (1024, 768, ["spawn"])]
for W, H, idxs in plan:
    game = orc.Game(path, "E1M1", W, H)
    items = []
    for idx in idxs:
        x, y, a = game.player_start() if idx == "spawn" else (float(t) for t in views[idx])
        variants = [dict()] if (W, idx) != (320, 511) else [dict(), dict(phases=1), dict(phases=2), dict(phases=4), dict(timestamp=0.4)]
        for var in variants:
            img = game.render(x, y, a, timestamp=var.get("timestamp", 0.0), phases=var.get("phases", 7))
            items.append(dict(index=idx, crc32=zlib.crc32(img.tobytes()), checksum=str(drr.checksum_numpy(img)), **var))
    out["frames"].append(dict(W=W, H=H, views=items))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_frames.json"), "w"), indent=1)
print("wrote", sum(len(c["views"]) for c in out["frames"]), "golden frames")
