"""Shared helpers for the test-suite (CPU and GPU)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from doom_rust_renderer_b200 import lib as drr  # noqa: E402
drr.use_test_library()  # the suite (and the processes it spawns) runs against libdrr_test.so: product objects + drr_test_* accessors
from doom_rust_renderer_b200 import synth_wad  # noqa: E402
from oracle import orc  # noqa: E402

CACHE = os.environ.get("DRR_CACHE", os.path.join(ROOT, "tests", "_cache"))
_MAPS = {}


def wad(kind: str = "e1m1"):
    """(path, GridMap) of the deterministic synthetic IWAD `kind`, generated on first use."""
    if kind not in _MAPS:
        os.makedirs(CACHE, exist_ok=True)
        data, gm, stats = synth_wad.build_wad(kind)
        path = os.path.join(CACHE, "synth_%s.wad" % kind)
        if not os.path.exists(path) or open(path, "rb").read() != data:
            with open(path + ".tmp", "wb") as f:
                f.write(data)
            os.replace(path + ".tmp", path)
        _MAPS[kind] = (path, gm, stats)
    return _MAPS[kind][0], _MAPS[kind][1]


def usable_views(game: orc.Game, views: np.ndarray, limit: int | None = None) -> np.ndarray:
    """Drop viewpoints on which the reference itself would panic (e.g. a seg through the eye point)."""
    keep = []
    for v in views:
        try:
            game.render(float(v[0]), float(v[1]), float(v[2]))
            keep.append(v)
        except orc.OracleError:
            pass
        if limit and len(keep) >= limit:
            break
    return np.array(keep, np.float32).reshape(-1, 3)


def upload_oracle_assets(ctx: drr.Context, game: orc.Game):
    """Upload the oracle's own decoded assets under the oracle's ids (for trace-replay tests)."""
    ctx.upload_palette(game.palette())
    for i in range(game.bitmap_count()):
        bm = game.bitmap(i)
        if bm.shape[0] > 0 and bm.shape[1] > 0:
            ctx.upload_bitmap(i, bm)
    for i in range(game.flat_count()):
        ctx.upload_flat(i, game.flat(i))
    ctx.set_sky(game.sky_bitmap_id())


def emit_trace(ctx: drr.Context, game: orc.Game, view_idx: int, x: float, y: float, angle: float, trace, phases=7):
    """Feed the oracle's recorded leaf calls (render(trace=True)) through the C ABI, in call order."""
    import math
    fh = game.floor_height_at(x, y)
    ang = np.float32(angle)
    # cos/sin must come from the same libm as the oracle: take them from numpy float32 -> C cosf? No: ask the oracle's
    # libc directly so there is no doubt.
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float
    libm.cosf.argtypes = [ctypes.c_float]
    libm.sinf.restype = ctypes.c_float
    libm.sinf.argtypes = [ctypes.c_float]
    ctx.frame_begin(view_idx, x, y, fh, float(ang), libm.cosf(float(ang)), libm.sinf(float(ang)))
    i = 0
    n = len(trace)
    while i < n:
        t = trace[i]
        if t["kind"] == 1:
            if phases & 2:
                hdr = drr.DrrVisplaneHdr(drr.FLAT_SKY if t["is_sky"] else t["asset"], t["height"], t["light_level"], t["left"], t["right"], 0)
                ctx.emit_visplane(hdr, t["top"][t["left"]:t["right"] + 1], t["bottom"][t["left"]:t["right"] + 1])
            i += 1
            continue
        # group consecutive column calls that share the per-seg arguments
        j = i
        key = (t["phase"], t["asset"], t["light_level"], tuple(t["line"]), float(t["start_offset"]), t["start_x"], t["end_x"],
               float(t["bottom_height"]), float(t["top_height"]), t["offset_x"], t["offset_y"])
        cols = []
        while j < n and trace[j]["kind"] == 0:
            u = trace[j]
            k2 = (u["phase"], u["asset"], u["light_level"], tuple(u["line"]), float(u["start_offset"]), u["start_x"], u["end_x"],
                  float(u["bottom_height"]), float(u["top_height"]), u["offset_x"], u["offset_y"])
            if k2 != key:
                break
            cols.append((u["x"], u["clipped_top_y"], u["clipped_bottom_y"], u["bottom_y"], u["top_y"]))
            j += 1
        want = (phases & 1) if t["phase"] == 0 else (phases & 4)
        if want:
            hdr = drr.DrrSegHdr(t["asset"], t["light_level"], t["phase"], t["line"][0], t["line"][1], t["line"][2], t["line"][3],
                                t["start_offset"], t["start_x"], t["end_x"], t["bottom_height"], t["top_height"], t["offset_x"], t["offset_y"])
            ctx.emit_columns(hdr, np.array(cols, dtype=drr.COL_DTYPE))
        i = j
    ctx.frame_end()


class AssetsFromCtx:
    """Texel data read back from a context's host mirrors (test accessors), keyed by device slot."""

    def __init__(self, ctx: drr.Context):
        import ctypes as C
        self.ctx = ctx
        self._bm = {}
        self._fl = {}
        pal = np.zeros(768, np.uint8)
        assert ctx.L.drr_test_palette(ctx.h, pal.ctypes.data_as(C.c_void_p)) == 0
        self.palette = pal
        self.sky_slot = ctx.L.drr_test_sky_slot(ctx.h)

    def bitmap(self, slot: int) -> np.ndarray:
        import ctypes as C
        if slot not in self._bm:
            w, h, o = C.c_int(), C.c_int(), C.c_int()
            assert self.ctx.L.drr_test_bitmap_info(self.ctx.h, slot, C.byref(w), C.byref(h), C.byref(o)) == 0
            a = np.zeros((h.value, w.value), np.int16)
            assert self.ctx.L.drr_test_bitmap_texels(self.ctx.h, slot, a.ctypes.data_as(C.c_void_p)) == 0
            self._bm[slot] = a
        return self._bm[slot]

    def flat(self, slot: int) -> np.ndarray:
        import ctypes as C
        if slot not in self._fl:
            a = np.zeros(4096, np.uint8)
            assert self.ctx.L.drr_test_flat_texels(self.ctx.h, slot, a.ctypes.data_as(C.c_void_p)) == 0
            self._fl[slot] = a
        return self._fl[slot]


def replay_binned_frame(ctx: drr.Context, frame: int) -> np.ndarray:
    """CPU replay of ONE recorded frame from the column-binned lists (the host restatement of what the bin kernel builds),
    span by span in draw order, using the oracle's leaf drawers.  Validates recording + binning without a GPU."""
    W, H = ctx.W, ctx.H
    assets = AssetsFromCtx(ctx)
    leaf = orc.Leaf(W, H, assets.palette)
    views = ctx._list(0, drr.VIEW_DTYPE)
    segs = ctx._list(1, drr.SEG_DTYPE)
    planes = ctx._list(2, drr.PLANE_DTYPE)
    spans = ctx._list(3, drr.SPAN_DTYPE)
    colidx = ctx._list(4, drr.COLIDX_DTYPE)
    v = views[frame]
    img = np.zeros((H, W, 3), np.uint8)
    scratch = np.zeros((H, W, 3), np.uint8)
    sky = assets.bitmap(assets.sky_slot) if assets.sky_slot >= 0 else None
    top = np.zeros(W, np.int16)
    bottom = np.zeros(W, np.int16)
    for x in range(W):
        ci = colidx[frame * W + x]
        lst = spans[ci["first"]:ci["first"] + ci["n"]]  # the column's spans in draw order
        for s in lst:
            assert s["x"] == x
            y0, y1 = int(s["y0"]), int(s["y1"])
            if s["kind"] in (drr.KIND_WALL, drr.KIND_WALL_HOLES):
                g = segs[s["op"]]
                leaf.column(img, assets.bitmap(int(g["bitmap_slot"])), int(g["light_level"]), (g["lsx"], g["lsy"], g["lex"], g["ley"]),
                            g["start_offset"], int(g["start_x"]), int(g["end_x"]), g["bottom_height"], g["top_height"], int(g["offset_x"]),
                            int(g["offset_y"]), x, y1, y0, int(s["bottom_y"]), int(s["top_y"]))
            else:
                p = planes[s["op"]]
                is_sky = s["kind"] in (drr.KIND_SKY, drr.KIND_SKY_HOLES)
                # the oracle's flat drawer skips columns with bottom-top <= 1; a resolved span may be that short, so draw a
                # padded range into a scratch image and copy the rows the span owns
                a0, a1 = (y0, y1) if is_sky else (max(0, y0 - 2), min(H - 1, y1 + 2))
                top[x], bottom[x] = a0, a1
                scratch[a0:a1 + 1, x] = img[a0:a1 + 1, x]
                leaf.visplane(scratch, None if is_sky else assets.flat(int(p["flat_slot"])), sky if is_sky else None, top, bottom,
                              1 if is_sky else 0, int(p["height"]), int(p["light_level"]), x, x, v["pos_x"], v["pos_y"], v["floor_height"], v["angle"])
                img[y0:y1 + 1, x] = scratch[y0:y1 + 1, x]
    return img
