"""CPU tests of the oracle itself.

The reference ships no tests, golden vectors or fixtures for this path, so the oracle is "parity
unpinned" by the reference; what can be pinned is (a) the Rust scalar semantics it relies on, checked here against
independent numpy/python restatements, and (b) its own output on the deterministic synthetic WAD, frozen as golden
checksums in tests/golden/ (regenerate with tools/make_golden.py)."""
from __future__ import annotations

import json
import os
import zlib

import numpy as np
import pytest

import common
from common import drr, orc, synth_wad

GOLDEN = os.path.join(common.ROOT, "tests", "golden", "oracle_frames.json")


def test_synthetic_wad_is_deterministic():
    data, gm, stats = synth_wad.build_wad("e1m1")
    data2, _, _ = synth_wad.build_wad("e1m1")
    assert data == data2
    gold = json.load(open(GOLDEN))
    assert zlib.crc32(data) == gold["wad_crc32"]["e1m1"]
    assert stats["sectors"] >= 85 and stats["segs"] >= 730 and stats["things"] >= 130


def test_oracle_frames_match_golden():
    gold = json.load(open(GOLDEN))
    path, gm = common.wad("e1m1")
    views = synth_wad.walk_viewpoints(gm, 4096)
    for case in gold["frames"]:
        W, H = case["W"], case["H"]
        game = orc.Game(path, "E1M1", W, H)
        for item in case["views"]:
            if item["index"] == "spawn":
                x, y, a = game.player_start()
            else:
                x, y, a = (float(t) for t in views[item["index"]])
            img = game.render(x, y, a, timestamp=item.get("timestamp", 0.0), phases=item.get("phases", 7))
            assert zlib.crc32(img.tobytes()) == item["crc32"], (W, H, item)
            assert drr.checksum_numpy(img) == int(item["checksum"]), (W, H, item)


def _f32(x):
    return np.float32(x)


def test_diminish_color_against_numpy_float32():
    """bitmap_render.rs:190-208 restated independently with numpy float32 scalars."""
    rng = np.random.default_rng(3)
    for _ in range(4000):
        rgb = [int(v) for v in rng.integers(0, 256, 3)]
        light = int(rng.integers(-50, 400))
        dist = int(rng.choice([rng.integers(-32768, 32768), rng.integers(-200, 3000)]))
        factor = _f32(light) / _f32(255.0)
        factor = _f32(factor - _f32(_f32(dist) * _f32(1.0 / 4096.0)))
        if factor < 0:
            factor = _f32(0)
        want = []
        for c in rgb:
            v = _f32(_f32(c) * factor)
            want.append(0 if not (v > 0) else 255 if v >= 255 else int(v))
        assert orc.diminish_color(rgb, light, dist) == tuple(want)


def test_wrap_mod_idiom_is_floormod():
    """`if t < 0 { t += n * (1 - t / n) } t %= n` in wrapping i16 equals the mathematical floor-mod for every i16 t and
    every n in 1..=32767 the kernels can meet (the device code relies on this identity, drr_kernels.cu: wall_ty)."""
    t = np.arange(-32768, 32768, dtype=np.int64)
    for n in list(range(1, 300)) + [511, 512, 1000, 4096, 30000, 32767]:
        q = np.trunc(t / n).astype(np.int64)  # i16 division truncates toward zero
        one_minus = ((1 - q + 32768) % 65536) - 32768
        prod = ((n * one_minus + 32768) % 65536) - 32768
        fixed = np.where(t < 0, ((t + prod + 32768) % 65536) - 32768, t)
        res = np.fmod(fixed, n).astype(np.int64)  # remainder takes the dividend's sign
        assert (res == np.mod(t, n)).all(), n


def test_leaf_column_matches_python_restatement():
    """render_vertical_bitmap_line restated in python/numpy float32 for one seg, compared pixel by pixel."""
    W, H = 64, 48
    rng = np.random.default_rng(11)
    pal = rng.integers(0, 256, 768, dtype=np.uint8)
    tex = rng.integers(0, 256, (72, 40)).astype(np.int16)
    tex[rng.random(tex.shape) < 0.2] = -1
    leaf = orc.Leaf(W, H, pal)
    line = (np.float32(100.0), np.float32(40.0), np.float32(180.5), np.float32(-30.25))
    so, bh, th = np.float32(12.5), np.float32(-41.0), np.float32(87.0)
    sx, ex, light, ox, oy = 3, 60, 200, 17, -9
    img = np.zeros((H, W, 3), np.uint8)
    want = np.zeros((H, W, 3), np.uint8)
    f = np.float32
    for x in range(sx, ex + 1):
        top_y, bottom_y = 5 - x // 8, 40 + x // 6
        ct, cb = max(top_y, 0), min(bottom_y, H - 1)
        leaf.column(img, tex, light, line, so, sx, ex, bh, th, ox, oy, x, cb, ct, bottom_y, top_y)
        ln = f(np.sqrt(f(f(f(line[0] - line[2]) ** 2) + f(f(line[1] - line[3]) ** 2))))
        ax = f(f(x - sx) / f(ex - sx))
        num = f(f(f(1 - ax) * f(f(0) / line[0])) + f(ax * f(ln / line[2])))
        den = f(f(f(1 - ax) * f(f(1) / line[0])) + f(ax * f(f(1) / line[2])))
        tx = int(f(num / den)) + int(so) + ox
        tx %= tex.shape[1]
        z = int(f(f(f(1 - ax) + ax) / den))
        fac = f(f(f(light) / f(255)) - f(f(z) * f(1 / 4096)))
        fac = f(max(fac, f(0)))
        for y in range(ct, cb + 1):
            ay = f(f(y - top_y) / f(bottom_y - top_y))
            ty = int(f(f(tex.shape[0]) + f(ay * f(th - bh)))) + oy
            ty %= tex.shape[0]
            t = int(tex[ty, tx])
            if t >= 0:
                want[y, x] = [min(255, int(f(f(pal[3 * t + c]) * fac))) for c in range(3)]
    assert (img == want).all()


def test_leaf_visplane_matches_python_restatement():
    W, H = 64, 48
    rng = np.random.default_rng(12)
    pal = rng.integers(0, 256, 768, dtype=np.uint8)
    flat = rng.integers(0, 256, 4096, dtype=np.uint8)
    leaf = orc.Leaf(W, H, pal)
    f = np.float32
    px, py, fh, ang = f(-1503.5), f(977.25), f(32.0), f(0.6)
    top = np.full(W, 30, np.int16)
    bottom = np.full(W, 47, np.int16)
    img = np.zeros((H, W, 3), np.uint8)
    leaf.visplane(img, flat, None, top, bottom, 0, 32, 180, 2, 61, px, py, fh, ang)
    import ctypes
    m = ctypes.CDLL("libm.so.6")
    m.cosf.restype = m.sinf.restype = ctypes.c_float
    m.cosf.argtypes = m.sinf.argtypes = [ctypes.c_float]
    ca, sa = f(m.cosf(float(ang))), f(m.sinf(float(ang)))
    aspect = f(f(200.0) / f(240.0))
    gcfx = f(f(f(W) / aspect) / f(2))
    want = np.zeros((H, W, 3), np.uint8)

    def i16(v):
        v = float(v)
        return 0 if v != v else max(-32768, min(32767, int(v)))
    for x in range(2, 62):
        for y in range(30, 48):
            vx = f(f(f(W / 2) - f(x)) / aspect)
            vy = f(f(H / 2) - f(y))
            wz = f(f(f(32) - fh) - f(41))
            with np.errstate(divide="ignore", invalid="ignore"):
                wx = f(f(gcfx * wz) / vy)
                wy = f(f(wz * vx) / vy)
            rx = f(f(wx * ca) - f(wy * sa))
            ry = f(f(wy * ca) + f(wx * sa))
            tx = (i16(rx) + i16(px)) & 63
            ty = (i16(ry) + i16(py)) & 63
            fac = f(f(f(180) / f(255)) - f(f(i16(wx)) * f(1 / 4096)))
            fac = f(max(fac, f(0)))
            t = int(flat[ty * 64 + tx])
            want[y, x] = [min(255, int(f(f(pal[3 * t + c]) * fac))) for c in range(3)]
    assert (img == want).all()


def test_trace_replay_closure_on_cpu():
    """direct render == the recorded leaf calls replayed in order (oracle self-consistency)."""
    path, gm = common.wad("e1m1")
    W, H = 160, 100
    game = orc.Game(path, "E1M1", W, H)
    v = common.usable_views(game, synth_wad.walk_viewpoints(gm, 64)[5::9], 3)
    for x, y, a in v:
        ref = game.render(float(x), float(y), float(a), trace=True)
        tr = game.trace()
        leaf = orc.Leaf(W, H, game.palette())
        img = np.zeros((H, W, 3), np.uint8)
        fh = game.floor_height_at(float(x), float(y))
        sky = game.bitmap(game.sky_bitmap_id())
        for t in tr:
            if t["kind"] == 0:
                leaf.column(img, game.bitmap(t["asset"]), t["light_level"], t["line"], t["start_offset"], t["start_x"], t["end_x"],
                            t["bottom_height"], t["top_height"], t["offset_x"], t["offset_y"], t["x"], t["clipped_bottom_y"],
                            t["clipped_top_y"], t["bottom_y"], t["top_y"])
            else:
                leaf.visplane(img, game.flat(t["asset"]), sky, t["top"], t["bottom"], t["is_sky"], t["height"], t["light_level"],
                              t["left"], t["right"], x, y, fh, a)
        assert (img == ref).all()


# ---- hand-derived known-answer tests (the derivation is in each docstring; nothing here is computed by code under test) ------
def test_kat_diminish_color():
    """bitmap_render.rs:190-208.  factor = light/255 - distance/4096, clamped below at 0 only; `as u8` saturates at 255.
      (200,100,50), light 300, dist 0     : factor 1.17647   -> 235.29, 117.65, 58.82  -> (235, 117, 58)
      (200,100,50), light 400, dist 0     : factor 1.56863   -> 313.7 (saturates), 156.86, 78.43 -> (255, 156, 78)
      (200,100,50), light 255, dist -4096 : factor 1 + 1 = 2 -> 400 (saturates), 200, 100 -> (255, 200, 100)
      (200,100,50), light 128, dist 1024  : factor 0.50196 - 0.25 = 0.25196 -> 50.39, 25.19, 12.59 -> (50, 25, 12)
      (200,100,50), light 0,   dist 100   : factor -0.0244 -> 0 -> (0, 0, 0)
      (255,255,255), light 255, dist 0    : factor exactly 1 -> (255, 255, 255)"""
    assert orc.diminish_color((200, 100, 50), 300, 0) == (235, 117, 58)
    assert orc.diminish_color((200, 100, 50), 400, 0) == (255, 156, 78)
    assert orc.diminish_color((200, 100, 50), 255, -4096) == (255, 200, 100)
    assert orc.diminish_color((200, 100, 50), 128, 1024) == (50, 25, 12)
    assert orc.diminish_color((200, 100, 50), 0, 100) == (0, 0, 0)
    assert orc.diminish_color((255, 255, 255), 255, 0) == (255, 255, 255)


def _grey_palette():
    return np.repeat(np.arange(256, dtype=np.uint8), 3)  # colour i = (i, i, i)


def test_kat_sky_texture_column_for_negative_and_large_angles():
    """visplanes.rs:54-66 at 320x200, screen column x = 100: tx = (100 * 256 / 320) as i16 = 80, then (80 + tx_offset) % 256 with
    tx_offset = (-256 * angle / (pi/2)) as i16 + 256, wrapped up by whole textures when negative:
      angle -pi/4 : -256 * -0.5 = 128          -> 128 + 256 = 384                                   -> (80 + 384) % 256 = 208
      angle  pi   : -256 * 2 = -512            -> -256 < 0 -> += 256 * (1 - (-256 / 256)) = +512    -> 256 -> (80 + 256) % 256 = 80
      angle 3pi/4 : -256 * 1.5 = -384          -> -128 < 0 -> += 256 * (1 - (-128 / 256 = 0)) = +256 -> 128 -> (80 + 128) % 256 = 208
      angle  0    : 0                          -> 256                                               -> (80 + 256) % 256 = 80
    (f32(pi)/4, f32(pi)/2 and f32(pi) are exact binary multiples of each other, so the quotients are exact.)
    Row y = 150: ty = (150 * 128 * 2 / 200) as i16 = 192, % 128 = 64.  The sky texture encodes tx in one test and ty in the other."""
    W, H = 320, 200
    leaf = orc.Leaf(W, H, _grey_palette())
    tex_tx = np.tile(np.arange(256, dtype=np.int16), (128, 1))            # texel = tx
    tex_ty = np.tile(np.arange(128, dtype=np.int16)[:, None], (1, 256))   # texel = ty
    top, bottom = np.full(W, 150, np.int16), np.full(W, 150, np.int16)
    pi = np.float32(np.pi)
    for angle, want_tx in ((-pi / 4, 208), (pi, 80), (3 * pi / 4, 208), (np.float32(0), 80)):
        img = np.zeros((H, W, 3), np.uint8)
        leaf.visplane(img, None, tex_tx, top, bottom, 1, 0, 255, 100, 100, 0.0, 0.0, 0.0, angle)
        assert img[150, 100].tolist() == [want_tx] * 3, (float(angle), img[150, 100])
        assert img.sum() == 3 * want_tx  # nothing else was drawn
    img = np.zeros((H, W, 3), np.uint8)
    leaf.visplane(img, None, tex_ty, top, bottom, 1, 0, 255, 100, 100, 0.0, 0.0, 0.0, 0.0)
    assert img[150, 100].tolist() == [64] * 3


def test_kat_texture_row_of_a_zero_height_column():
    """bitmap_render.rs:253-265 when bottom_y == top_y (a zero-height sector, e.g. sector 16 of E1M1): ay = (y - top_y) / 0 is NaN
    (y == top_y) or +-inf, (1 - ay) * 0.0 is NaN either way, the sum is NaN and `NaN as i16` is 0; so every drawn row takes
    texture row (0 + offset_y) mod height.  Texture 16 rows high whose texel = its row, offset_y = 5 -> palette entry 5;
    offset_y = -3 -> ty = -3 < 0 -> += 16 * (1 - (-3 / 16 = 0)) = 13 -> entry 13.  Light 255 at depth 0 -> factor 1."""
    W, H = 64, 48
    leaf = orc.Leaf(W, H, _grey_palette())
    tex = np.tile(np.arange(16, dtype=np.int16)[:, None], (1, 8))
    line = (np.float32(0.0), np.float32(10.0), np.float32(0.0), np.float32(-10.0))  # uz0 = uz1 = 0: z = (1 / (0/0 ...)) -> NaN -> 0
    for oy, want in ((5, 5), (-3, 13)):
        img = np.zeros((H, W, 3), np.uint8)
        leaf.column(img, tex, 255, line, np.float32(0.0), 0, 10, np.float32(0.0), np.float32(64.0), 0, oy, 5, 22, 18, 20, 20)
        assert [img[y, 5, 0] for y in range(18, 23)] == [want] * 5, (oy, img[16:25, 5, 0])


# ---- the front-end's stateless half, pinned by an independent restatement written from the Rust ---------------------------------
@pytest.mark.parametrize("kind,W,H,step", [("e1m1", 320, 200, 409), ("e1m1", 1280, 800, 1021), ("stress", 640, 400, 3)])
def test_oracle_wall_calls_carry_the_arguments_an_independent_restatement_derives(kind, W, H, step):
    """tests/ref_frontend.py restates, from src/renderer/misc.rs, segs.rs and src/map/*.rs, how a seg becomes the arguments of
    render_vertical_bitmap_line (own WAD reader, view transform, clip_to_viewport, make_sidedef_non_vertical_line, the parts
    of process_seg, the per-column bottom_y / top_y).  Every immediate wall call in the oracle's trace must carry, bit for
    bit, the arguments derived there for some part of some seg, and its rows must obey process_sidedef's clipping rules."""
    import ref_frontend as rf
    path, gm = common.wad(kind)
    game = orc.Game(path, "E1M1", W, H)
    m = rf.MapLumps(path)
    k = rf.Constants(W, H)
    allv = synth_wad.walk_viewpoints(gm, 4096)[::step] if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, 64)[::step]
    views = common.usable_views(game, allv, 10 if (kind, W) == ("e1m1", 320) else 4)
    checked = masked = 0
    bits = lambda v: np.float32(v).view(np.uint32).item()  # noqa: E731
    for v in views:
        x, y, a = (np.float32(t) for t in v)
        game.render(float(x), float(y), float(a), trace=True)
        fh = game.floor_height_at(float(x), float(y))
        cands = {}
        for seg in m.segs:
            for p in rf.seg_parts(k, m, seg, x, y, a, fh):
                key = tuple(bits(t) for t in p["line"]) + (bits(p["start_offset"]), p["start_x"], p["end_x"], bits(p["bottom_height"]),
                                                           bits(p["top_height"]), p["offset_x"], p["offset_y"], p["light_level"])
                cands[key] = p
        for t in game.trace():
            if t["kind"] != 0:
                continue
            key = tuple(bits(q) for q in t["line"]) + (bits(t["start_offset"]), t["start_x"], t["end_x"], bits(t["bottom_height"]),
                                                       bits(t["top_height"]), t["offset_x"], t["offset_y"], t["light_level"])
            p = cands.get(key)
            if t["phase"] != 0:  # drawn late: a masked mid-texture (must be a two-sided middle part) or a sprite (not a seg at all)
                masked += p is not None and p["flags"] == "two_sided_middle"
                continue
            assert p is not None, ("no seg part yields this wall call", v, t)
            assert p["flags"] in ("solid", "lower", "upper"), p["flags"]
            assert p["start_x"] <= t["x"] <= p["end_x"]
            assert p["column"](t["x"]) == (t["bottom_y"], t["top_y"]), (v, t, p["column"](t["x"]))
            # segs.rs:189-200: clipped to the occlusion arrays and the screen, drawn only when not empty
            assert t["clipped_bottom_y"] <= min(H - 1, t["bottom_y"]) and t["clipped_top_y"] >= max(0, t["top_y"])
            assert t["clipped_bottom_y"] >= t["clipped_top_y"]
            checked += 1
    assert checked > 50 * len(views), checked
    if (kind, W) == ("e1m1", 320):
        assert masked > 0  # (one of these viewpoints looks through the grates)


def test_oracle_late_draw_order_obeys_the_reference_rules():
    """The order of what is drawn late (map objects and masked mid-textures), checked against rules read off the Rust, not off
    the C++ oracle: draw_map_objects sorts the sprites by `clipped_line.line.start.x as i16` and reverses (map_objects.rs:216-217,
    Ord for BitmapRender bitmap_render.rs:167-173), and before a sprite is drawn every not yet drawn two-sided seg for which
    is_behind_vertex(midpoint of the sprite's line) holds is drawn (map_objects.rs:219-239, bitmap_render.rs:137-165).  So in
    the oracle's trace (1) the sprites' keys never increase, and (2) whenever a masked mid-texture M is behind the midpoint of a
    sprite S, M's columns come before S's.  Masked mid-textures are told from sprites by tests/ref_frontend.py (a late call
    that carries the arguments of a two-sided middle part of some seg)."""
    import ref_frontend as rf
    W, H = 320, 200
    path, gm = common.wad("e1m1")
    game = orc.Game(path, "E1M1", W, H)
    m = rf.MapLumps(path)
    k = rf.Constants(W, H)
    f32 = np.float32
    bits = lambda v: np.float32(v).view(np.uint32).item()  # noqa: E731

    def as_i16(v):  # Rust `as i16`
        v = float(v)
        return 0 if v != v else int(max(-32768.0, min(32767.0, np.trunc(v))))

    def is_behind_vertex(line, v):  # bitmap_render.rs:137-165, line = (sx, sy, ex, ey) in view space
        sx, sy, ex, ey = (f32(t) for t in line)
        min_x, max_x = min(sx, ex), max(sx, ex)
        if min_x > v[0]:
            return True
        return bool(max_x > v[0] and not rf.is_left_of_line(v, ((sx, sy), (ex, ey))))

    allv = synth_wad.walk_viewpoints(gm, 4096)
    # (the stretches of the walk from which the map's grates -- its masked mid-textures -- are in view, and a sample of the rest)
    views = common.usable_views(game, np.concatenate([allv[3320:3440:3], allv[3560:3704:3], allv[::173]]), 110)
    pairs = behind_pairs = sprites_seen = mids_seen = 0
    for v in views:
        x, y, a = (f32(t) for t in v)
        game.render(float(x), float(y), float(a), trace=True)
        fh = game.floor_height_at(float(x), float(y))
        mids_args = set()
        for seg in m.segs:
            for p in rf.seg_parts(k, m, seg, x, y, a, fh):
                if p["flags"] == "two_sided_middle":
                    mids_args.add(tuple(bits(t) for t in p["line"]) + (bits(p["start_offset"]), p["start_x"], p["end_x"], bits(p["bottom_height"]),
                                                                       bits(p["top_height"]), p["offset_x"], p["offset_y"], p["light_level"]))
        late = []  # BitmapRenders in draw order: (is_mid, line)
        last = None
        for t in game.trace():
            if t["kind"] != 0 or t["phase"] != 2:
                continue
            key = tuple(bits(q) for q in t["line"]) + (bits(t["start_offset"]), t["start_x"], t["end_x"], bits(t["bottom_height"]),
                                                       bits(t["top_height"]), t["offset_x"], t["offset_y"], t["light_level"])
            if (key, t["asset"]) != last:  # a BitmapRender draws its columns in one go (bitmap_render.rs:107-128)
                late.append((key in mids_args, tuple(f32(q) for q in t["line"])))
                last = (key, t["asset"])
        spr = [(i, line) for i, (is_mid, line) in enumerate(late) if not is_mid]
        mid = [(i, line) for i, (is_mid, line) in enumerate(late) if is_mid]
        sprites_seen += len(spr)
        mids_seen += len(mid)
        keys = [as_i16(line[0]) for _, line in spr]
        assert all(keys[i] >= keys[i + 1] for i in range(len(keys) - 1)), ("sprites not in descending key order", v, keys)
        for si, sl in spr:
            mp = (f32(f32(sl[0] + sl[2]) / f32(2.0)), f32(f32(sl[1] + sl[3]) / f32(2.0)))
            for mi, ml in mid:
                pairs += 1
                if is_behind_vertex(ml, mp):
                    behind_pairs += 1
                    assert mi < si, ("a masked mid-texture behind a sprite was drawn after it", v, ml, sl)
    print("late-draw order: %d sprites, %d masked mid-textures, %d pairs, %d of them \"behind\"" % (sprites_seen, mids_seen, pairs, behind_pairs))
    assert sprites_seen > 100 and mids_seen > 10 and behind_pairs > 10, (sprites_seen, mids_seen, pairs, behind_pairs)
