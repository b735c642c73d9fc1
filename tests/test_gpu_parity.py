"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.  Bit-exact (byte for byte) RGB24.

Tolerance: none.  The path is f32 + integer work whose every rounding is specified, so the bar
is equality of every byte of every frame, plus equality of the device checksum with the host checksum of the oracle frame.
"""
from __future__ import annotations

import ctypes

import numpy as np
import pytest

import common
from common import drr, orc, synth_wad

pytestmark = pytest.mark.gpu


def _libm():
    m = ctypes.CDLL("libm.so.6")
    for f in (m.cosf, m.sinf):
        f.restype = ctypes.c_float
        f.argtypes = [ctypes.c_float]
    return m


def _compare(ctx, k, ref, what):
    got = ctx.read_framebuffer(k)
    diff = (got != ref).any(2)
    if diff.any():
        ys, xs = np.nonzero(diff)
        msg = "%s: %d pixels differ; first at x=%d y=%d got=%s want=%s" % (what, int(diff.sum()), xs[0], ys[0], got[ys[0], xs[0]], ref[ys[0], xs[0]])
        raise AssertionError(msg)


# ---- whole frames through the product front-end -----------------------------------------------------------------------
# (1000x900: three row bands + a width that is no multiple of the tile; 96x2000: five bands; 64x3300: nine bands, more than the
# bin kernel keeps separate span lists for; 200x120, 324x200: bytewise store path + checksum pass)
@pytest.mark.parametrize("W,H,n", [(320, 200, 48), (640, 400, 12), (1280, 800, 6), (1024, 768, 3), (1920, 1200, 2), (200, 120, 6), (324, 200, 3),
                                   (1000, 900, 2), (96, 2000, 2), (64, 3300, 2)])
def test_scene_frames_match_oracle(W, H, n):
    path, gm = common.wad("e1m1")
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[:: max(1, 4096 // (n + 4))], n)
    assert len(views) == n
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    ctx.submit()
    ctx.sync()
    crcs = ctx.read_checksums(0, n)
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]))
        _compare(ctx, k, ref, "%dx%d view %d %s" % (W, H, k, v))
        assert int(crcs[k]) == drr.checksum_numpy(ref)


@pytest.mark.parametrize("W,H,kind", [(320, 200, "e1m1"), (200, 120, "stress"), (1280, 800, "e1m1")])
def test_device_binning_equals_host_restatement(W, H, kind):
    """The bin kernel's per-column span lists (rows, kind, draw order) equal the host restatement of the binning rule."""
    path, gm = common.wad(kind)
    game = orc.Game(path, "E1M1", W, H)
    src = synth_wad.walk_viewpoints(gm, 512) if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, 64)
    views = common.usable_views(game, src[:: max(1, len(src) // 8)], 5)
    ctx = drr.Context(W, H, 0, len(views))
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    ctx.submit()
    ctx.sync()
    ci_dev, recs = ctx.device_bins(len(views))
    nbands, band_rows, nlists = ctx.tile_bands()
    spans, ci_host = ctx._list(3, drr.SPAN_DTYPE), ctx._list(4, drr.COLIDX_DTYPE)
    assert ci_dev.shape == (len(views), nlists, W) and len(ci_host) == len(views) * W
    assert int(ci_host["n"].sum()) == ctx.stats()["spans"] == len(spans)
    covered = (ci_dev["n"] >> 31).astype(bool)  # bin kernel's flag: the always-writing spans cover every row of the band
    ci_dev["n"] &= 0x7FFFFFFF
    assert kind != "e1m1" or covered.mean() > 0.5
    # every device list lies inside the record slots and lists do not overlap
    flat = ci_dev.reshape(-1)
    order = np.argsort(flat["first"], kind="stable")
    nz = order[flat["n"][order] > 0]
    ends = flat["first"][nz].astype(np.int64) + flat["n"][nz]
    assert (ends[:-1] <= flat["first"][nz][1:]).all() and ends[-1] <= len(recs)
    for i in np.nonzero(ci_host["n"])[0]:
        f, x = divmod(int(i), W)
        h_all = spans[ci_host["first"][i]:ci_host["first"][i] + ci_host["n"][i]]
        for b in range(nlists):
            lo, hi = (b * band_rows, min(H, (b + 1) * band_rows) - 1) if nlists > 1 else (0, H - 1)
            h = h_all[(h_all["y1"] >= lo) & (h_all["y0"] <= hi)]  # the spans that touch the band, in draw order
            c = ci_dev[f, b, x]
            assert c["n"] == len(h), (f, b, x)
            d = recs[c["first"]:c["first"] + c["n"]]
            assert ((d[:, 0] & 0xFFFF) == h["y0"]).all() and ((d[:, 0] >> 16) == h["y1"]).all(), (f, b, x)
            kind_d = d[:, 1] & 0xFF
            assert ((kind_d == h["kind"]) | (kind_d == 7)).all(), (f, b, x)  # 7 = column the reference would have panicked on
            rows = np.zeros(H, np.int32)
            for k, sp in zip(kind_d, h):
                if k in (drr.KIND_WALL, drr.KIND_FLAT, drr.KIND_SKY):
                    rows[sp["y0"]:sp["y1"] + 1] += 1
            if covered[f, b, x]:
                assert (rows[lo:hi + 1] >= 1).all(), (f, b, x)
            elif int(np.isin(h_all["kind"], (drr.KIND_WALL, drr.KIND_FLAT, drr.KIND_SKY)).sum()) <= 6:  # (it tracks 6 spans per column)
                assert not (rows[lo:hi + 1] >= 1).all(), (f, b, x)


def test_config1_spawn_viewpoint():
    """BASELINE config 1: the Player1Start viewpoint at 320x200, timestamp 0."""
    path, _ = common.wad("e1m1")
    game = orc.Game(path, "E1M1", 320, 200)
    x, y, a = game.player_start()
    ctx = drr.Context(320, 200, 0, 1)
    scene = drr.Scene(path, "E1M1", 320, 200)
    assert scene.player_start() == (x, y, a)
    scene.upload_assets(ctx)
    scene.emit_view(ctx, 0, x, y, a)
    ctx.submit()
    ctx.sync()
    _compare(ctx, 0, game.render(x, y, a), "spawn view")


@pytest.mark.parametrize("phases", [1, 2, 4, 3, 6])
def test_phase_isolation(phases):
    """BASELINE config 3: walls-only / flats-only (/ masked-only) isolation."""
    path, gm = common.wad("e1m1")
    W, H = 640, 400
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[100::700], 5)
    ctx = drr.Context(W, H, 0, len(views))
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views, phases=phases) == []
    ctx.submit()
    ctx.sync()
    for k, v in enumerate(views):
        _compare(ctx, k, game.render(float(v[0]), float(v[1]), float(v[2]), phases=phases), "phases=%d view %d" % (phases, k))


@pytest.mark.parametrize("timestamp", [0.0, 0.4, 0.7, 12.5])
def test_animated_flats(timestamp):
    path, gm = common.wad("e1m1")
    W, H = 320, 200
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[50::400], 10)
    ctx = drr.Context(W, H, 0, len(views))
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views, timestamp=timestamp) == []
    ctx.submit()
    ctx.sync()
    for k, v in enumerate(views):
        _compare(ctx, k, game.render(float(v[0]), float(v[1]), float(v[2]), timestamp=timestamp), "t=%g view %d" % (timestamp, k))


def test_stress_map():
    """BASELINE config 5's map (open sectors, many visplanes, tall columns), a few viewpoints."""
    path, gm = common.wad("stress")
    W, H = 640, 400
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.scatter_viewpoints(gm, 64), 6)
    ctx = drr.Context(W, H, 0, len(views))
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    ctx.submit()
    ctx.sync()
    for k, v in enumerate(views):
        _compare(ctx, k, game.render(float(v[0]), float(v[1]), float(v[2])), "stress view %d" % k)


# ---- kernels alone: the oracle's own leaf-call trace fed through the C ABI --------------------------------------------
def test_oracle_trace_through_c_abi():
    path, gm = common.wad("e1m1")
    W, H = 320, 200
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[7::500], 8)
    # render once so that every lazily composed texture exists before the assets are exported
    refs = [game.render(float(v[0]), float(v[1]), float(v[2])) for v in views]
    ctx = drr.Context(W, H, 0, len(views))
    common.upload_oracle_assets(ctx, game)
    for k, v in enumerate(views):
        game.render(float(v[0]), float(v[1]), float(v[2]), trace=True)
        common.emit_trace(ctx, game, k, float(v[0]), float(v[1]), float(v[2]), game.trace())
    ctx.submit()
    ctx.sync()
    for k in range(len(views)):
        _compare(ctx, k, refs[k], "trace view %d" % k)


# ---- fuzz: arbitrary (also nonsensical) leaf arguments -----------------------------------------------------------------
def _fuzz_assets(rng, n_bitmaps=10):
    pal = rng.integers(0, 256, 768, dtype=np.uint8)
    bitmaps = []
    for i in range(n_bitmaps):
        w = int(rng.choice([1, 3, 8, 16, 41, 64, 128, 200]))
        h = int(rng.choice([1, 2, 7, 16, 56, 72, 96, 128, 130]))
        t = rng.integers(0, 256, (h, w)).astype(np.int16)
        if i % 2:
            t[rng.random((h, w)) < 0.35] = -1
        bitmaps.append(t)
    sky = rng.integers(0, 256, (128, 256)).astype(np.int16)
    flats = [rng.integers(0, 256, 4096, dtype=np.uint8) for _ in range(4)]
    return pal, bitmaps, sky, flats


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("W,H", [(320, 200), (96, 64), (64, 900)])
def test_fuzz_columns(seed, W, H):
    """Random render_vertical_bitmap_line arguments, including degenerate ones (zero-height columns, bottom_y == top_y,
    huge offsets, NaN / inf line ends, negative light, transparent texels, overlapping columns in draw order)."""
    rng = np.random.default_rng(seed)
    pal, bitmaps, sky, flats = _fuzz_assets(rng)
    nviews = 4
    ctx = drr.Context(W, H, 0, nviews)
    ctx.upload_palette(pal)
    for i, b in enumerate(bitmaps):
        ctx.upload_bitmap(100 + i, b)
    ctx.upload_bitmap(99, sky)
    ctx.set_sky(99)
    leaf = orc.Leaf(W, H, pal)
    refs = []
    special = [np.float32(v) for v in (0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, 3e38)]
    for v in range(nviews):
        ref = np.zeros((H, W, 3), np.uint8)
        ctx.frame_begin(v, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)
        for _ in range(120):
            bi = int(rng.integers(0, len(bitmaps)))
            line = rng.uniform(0.5, 900, 4).astype(np.float32)
            line[1] = rng.uniform(-600, 600)
            line[3] = rng.uniform(-600, 600)
            so = np.float32(rng.uniform(0, 300))
            bh, th = np.float32(rng.uniform(-200, 50)), np.float32(rng.uniform(-50, 400))
            if rng.random() < 0.08:
                line[int(rng.integers(0, 4))] = rng.choice(special)
            if rng.random() < 0.05:
                bh = rng.choice(special)
            if rng.random() < 0.05:
                so = rng.choice(special)
            sx = int(rng.integers(-20, W))
            ex = sx + int(rng.integers(0, 80))
            light = int(rng.integers(-40, 300))
            ox, oy = int(rng.integers(-400, 400)), int(rng.integers(-400, 400))
            if rng.random() < 0.05:
                oy = int(rng.choice([-32768, 32767, -32000, 32000]))
            if rng.random() < 0.05:
                ox = int(rng.choice([-32768, 32767]))
            cols = []
            for x in range(max(sx, -2), min(ex + 1, W + 2)):
                if rng.random() < 0.3:
                    continue
                top_y = int(rng.integers(-300, H))
                bottom_y = top_y + int(rng.integers(0, 500)) if rng.random() > 0.05 else top_y
                ct = max(top_y, int(rng.integers(-5, H)))
                cb = min(bottom_y, int(rng.integers(0, H + 5)))
                cols.append((x, ct, cb, bottom_y, top_y))
            hdr = drr.DrrSegHdr(100 + bi, light, 0, line[0], line[1], line[2], line[3], so, sx, ex, bh, th, ox, oy)
            ctx.emit_columns(hdr, np.array(cols, dtype=drr.COL_DTYPE))
            for (x, ct, cb, by, ty) in cols:
                if 0 <= x < W:  # Pixels::set ignores x >= W; rows are clamped like the reference's y > H check
                    try:
                        leaf.column(ref, bitmaps[bi], light, line, so, sx, ex, bh, th, ox, oy, x, min(cb, H - 1), max(ct, 0), by, ty)
                    except orc.OracleError:
                        pytest.fail("oracle panicked on fuzz input (generator must avoid panics)")
        ctx.frame_end()
        refs.append(ref)
    ctx.submit()
    ctx.sync()
    for v in range(nviews):
        _compare(ctx, v, refs[v], "fuzz columns seed %d view %d" % (seed, v))


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("W,H", [(320, 200), (96, 64), (64, 900)])
def test_fuzz_visplanes(seed, W, H):
    """Random draw_visplane / draw_sky arguments: planes crossing the horizon (vy == 0 row, negative distances -> factor > 1),
    (0,0) columns (quirk Q3), 1-pixel columns (Q4), unclamped top/bottom, overlapping planes in draw order, any angle."""
    rng = np.random.default_rng(1000 + seed)
    pal, bitmaps, sky, flats = _fuzz_assets(rng, 1)
    if seed == 3:
        sky[rng.random(sky.shape) < 0.2] = -1  # sky with holes
    m = _libm()
    nviews = 4
    ctx = drr.Context(W, H, 0, nviews)
    ctx.upload_palette(pal)
    ctx.upload_bitmap(99, sky)
    ctx.set_sky(99)
    for i, f in enumerate(flats):
        ctx.upload_flat(10 + i, f)
    leaf = orc.Leaf(W, H, pal)
    refs = []
    for v in range(nviews):
        ref = np.zeros((H, W, 3), np.uint8)
        px, py = np.float32(rng.uniform(-3000, 3000)), np.float32(rng.uniform(-3000, 3000))
        fh = np.float32(rng.integers(-10, 10) * 8)
        ang = np.float32(rng.uniform(-7, 7))
        ctx.frame_begin(v, px, py, fh, ang, m.cosf(float(ang)), m.sinf(float(ang)))
        for _ in range(40):
            left = int(rng.integers(0, W))
            right = min(W - 1, left + int(rng.integers(0, 120)))
            top = np.zeros(W, np.int16)
            bottom = np.zeros(W, np.int16)
            t0 = int(rng.integers(-30, H))
            for x in range(left, right + 1):
                if rng.random() < 0.1:
                    continue  # stays (0, 0)
                t0 += int(rng.integers(-3, 4))
                top[x] = t0
                bottom[x] = t0 + int(rng.integers(-2, 90))
            is_sky = rng.random() < 0.25
            fi = int(rng.integers(0, len(flats)))
            height = int(rng.integers(-30, 60) * 8)
            light = int(rng.integers(0, 256))
            hdr = drr.DrrVisplaneHdr(drr.FLAT_SKY if is_sky else 10 + fi, height, light, left, right, 0)
            ctx.emit_visplane(hdr, top[left:right + 1], bottom[left:right + 1])
            leaf.visplane(ref, flats[fi], sky, top, bottom, 1 if is_sky else 0, height, light, left, right, px, py, fh, ang)
        ctx.frame_end()
        refs.append(ref)
    ctx.submit()
    ctx.sync()
    for v in range(nviews):
        _compare(ctx, v, refs[v], "fuzz visplanes seed %d view %d" % (seed, v))


def test_empty_and_black_frames():
    """A frame with no ops is all zeros (Pixels::new), and re-drawing a slot overwrites the previous contents."""
    W, H = 320, 200
    ctx = drr.Context(W, H, 0, 2)
    ctx.upload_palette(np.arange(768, dtype=np.uint8))
    for rep in range(2):
        ctx.reset()
        ctx.frame_begin(0, 0, 0, 0, 0, 1, 0)
        ctx.frame_end()
        ctx.frame_begin(1, 0, 0, 0, 0, 1, 0)
        ctx.frame_end()
        ctx.submit()
        ctx.sync()
        assert not ctx.read_framebuffer(0).any() and not ctx.read_framebuffer(1).any()
        assert list(ctx.read_checksums(0, 2)) == [0, 0]


def test_redraw_is_idempotent_and_checksum_of_checksums():
    path, gm = common.wad("e1m1")
    W, H = 320, 200
    n = 64
    views = synth_wad.walk_viewpoints(gm, n)
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    scene.emit_views(ctx, views)
    ctx.submit()
    ctx.sync()
    c1 = ctx.read_checksums(0, n)
    ctx.draw()
    ctx.sync()
    c2 = ctx.read_checksums(0, n)
    assert (c1 == c2).all()
    host = np.array([drr.checksum_numpy(ctx.read_framebuffer(k)) for k in range(n)], np.uint64)
    assert (host == c1).all()
    with np.errstate(over="ignore"):
        assert int(host.sum(dtype=np.uint64)) == int(c1.sum(dtype=np.uint64))


def test_pipelined_submit_equals_upload_then_draw(monkeypatch):
    """drr_submit cuts the batch into chunks of frames (copy stream + two compute streams); any chunking must give what
    drr_upload_lists + drr_draw gives, and drawing the same lists again after another batch reuses the buffers correctly."""
    path, gm = common.wad("e1m1")
    W, H, n = 320, 200, 23
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::150], n)
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    ctx.upload_lists()
    ctx.draw()
    ctx.sync()
    want = ctx.read_checksums(0, n)
    frame7 = ctx.read_framebuffer(7)
    assert int(want[7]) == drr.checksum_numpy(game.render(*[float(t) for t in views[7]]))
    for chunks in ("1", "2", "5", "23", "64"):
        ctx.set_knob("submit_chunks", int(chunks))
        ctx.submit()
        ctx.submit()  # back to back: the second upload must wait for the first draw
        ctx.sync()
        assert (ctx.read_checksums(0, n) == want).all(), chunks
        assert (ctx.read_framebuffer(7) == frame7).all(), chunks
    ctx.set_knob("submit_one_stream", 1)
    ctx.submit()
    assert (ctx.read_checksums(0, n) == want).all()


def test_full_size_batch_properties(monkeypatch):
    """BASELINE configs[1] at its full size (4096 viewpoints, 320x200), through properties that do not need 4096 oracle frames:
    the batch rendered whole == rendered as two halves in reversed slot order with another chunking (a frame's result depends
    on nothing but its own lists); a seeded sample of frames equals the oracle byte for byte; drawing twice changes nothing."""
    path, gm = common.wad("e1m1")
    W, H, n = 320, 200, 4096
    game = orc.Game(path, "E1M1", W, H)
    views = np.array(synth_wad.walk_viewpoints(gm, n), np.float32)
    scene = drr.Scene(path, "E1M1", W, H)
    ctx = drr.Context(W, H, 0, n)
    scene.upload_assets(ctx)
    skipped = set(scene.emit_views(ctx, views))  # viewpoints the reference itself would panic on keep an empty slot
    assert len(skipped) < n // 50
    ctx.submit()
    ctx.sync()
    whole = ctx.read_checksums(0, n)
    ctx.draw()
    ctx.sync()
    assert (ctx.read_checksums(0, n) == whole).all()
    rng = np.random.default_rng(0xD00D1993)
    for k in rng.choice(n, 24, replace=False):
        if int(k) in skipped:
            continue
        ref = game.render(float(views[k, 0]), float(views[k, 1]), float(views[k, 2]))
        _compare(ctx, int(k), ref, "full batch view %d" % k)
        assert int(whole[k]) == drr.checksum_numpy(ref)
    # two halves, slots reversed, one chunk per 64 frames
    monkeypatch.setenv("DRR_SUBMIT_CHUNKS", "7")
    half = n // 2
    ctx2 = drr.Context(W, H, 0, half)
    scene.upload_assets(ctx2)
    for lo in (0, half):
        ctx2.reset()
        for i in range(half):
            k = lo + i
            if k in skipped:
                continue
            scene.emit_view(ctx2, half - 1 - i, float(views[k, 0]), float(views[k, 1]), float(views[k, 2]))
        ctx2.submit()
        ctx2.sync()
        got = ctx2.read_checksums(0, half)[::-1]
        keep = np.array([lo + i not in skipped for i in range(half)])
        assert (got[keep] == whole[lo:lo + half][keep]).all()


def test_error_paths():
    ctx = drr.Context(64, 32, 0, 1)
    with pytest.raises(drr.DrrError):
        ctx.frame_end()  # not in a frame
    with pytest.raises(drr.DrrError):
        ctx.set_sky(5)  # unknown bitmap
    ctx.frame_begin(0, 0, 0, 0, 0, 1, 0)
    with pytest.raises(drr.DrrError):
        ctx.emit_columns(drr.DrrSegHdr(7, 0, 0, 1, 1, 1, 1, 0, 0, 1, 0, 1, 0, 0), np.zeros(0, drr.COL_DTYPE))  # unknown bitmap id
    ctx.frame_end()
    with pytest.raises(drr.DrrError):
        ctx.submit()  # palette missing


# ---- the hoisted-reciprocal division of the march kernel is the correctly rounded quotient ---------------------------------
def test_fast_division_walls_exhaustive():
    """bitmap_render.rs:256: ay = (y - top_y) as f32 / (bottom_y - top_y) as f32.  y in [0, H), top_y/bottom_y are i16, so the
    numerator is an integer in [-32767-H, 32768+H] and the denominator a non-zero integer in [-65535, 65535]: all
    ~1.8e10 pairs are compared bitwise with __fdiv_rn on the device."""
    ctx = drr.Context(64, 64, 0, 1)
    amax = 32768 + 2048
    bad, first = ctx.test_fastdiv(0, 2 * amax + 1, 2 * 65535)
    assert bad == 0, first


@pytest.mark.parametrize("H", [200, 400, 768, 800, 1200, 201])
def test_fast_division_flats_sampled(H):
    """visplanes.rs:113-114: wx = GCFX*wz / vy, wy = wz*vx / vy with vy = H/2 - y.  Every row's denominator against 2^25
    numerators spread over the whole float range (those outside [2^-60, 2^60] take __fdiv_rn in the kernel and are
    skipped here), including both signs."""
    ctx = drr.Context(64, 64, 0, 1)
    for lo in (0x00000000, 0x80000000, 0x3F800000 - (1 << 24), 0x12345678):
        bad, first = ctx.test_fastdiv(1, 1 << 23, H, H / 2.0, lo, 255)
        assert bad == 0, (H, lo, first)


# ---- contexts side by side: several devices in one process, large tiles in any order, recording while a batch uploads -----
def _device_count():
    import torch
    return torch.cuda.device_count()


def _draw_and_check(ctx, scene, game, views, what):
    assert scene.emit_views(ctx, views) == []
    ctx.submit()
    ctx.sync()
    crcs = ctx.read_checksums(0, len(views))
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]))
        _compare(ctx, k, ref, "%s view %d" % (what, k))
        assert int(crcs[k]) == drr.checksum_numpy(ref)


def test_two_devices_in_one_process():
    """drr.h: "distinct contexts may be driven from distinct threads, one per GPU".  Two contexts on two devices at 1280x800 (a
    tile needs more than the default 48 KB of dynamic shared memory: the opt-in is per device), driven interleaved from one
    thread -- every entry point binds its context's device -- and then from one thread each."""
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    import threading
    path, gm = common.wad("e1m1")
    W, H, n = 1280, 800, 3
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::331], 2 * n)
    scene = drr.Scene(path, "E1M1", W, H)
    ctxs = [drr.Context(W, H, d, n) for d in (0, 1)]
    for c in ctxs:
        scene.upload_assets(c)
    # one thread, calls interleaved
    for d, c in enumerate(ctxs):
        assert scene.emit_views(c, views[d * n:(d + 1) * n]) == []
    for c in ctxs:
        c.submit()
    for d, c in enumerate(ctxs):
        c.sync()
        for k in range(n):
            v = views[d * n + k]
            _compare(c, k, game.render(float(v[0]), float(v[1]), float(v[2])), "device %d view %d (one thread)" % (d, k))
    # one thread per context
    errors = []

    def work(d):
        try:
            c = ctxs[d]
            g = orc.Game(path, "E1M1", W, H)
            for rep in range(3):
                c.reset()
                _draw_and_check(c, drr.Scene(path, "E1M1", W, H), g, views[(1 - d) * n:(2 - d) * n], "device %d (own thread, pass %d)" % (d, rep))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]


def test_large_tiles_of_both_store_paths_in_one_process():
    """1280x800 (TMA write-out) and 1000x800 (width no multiple of 32: bytewise write-out), both with tiles above 48 KB, in
    either order on the same device: the shared-memory opt-in is kept per kernel instantiation."""
    path, gm = common.wad("e1m1")
    for order in ((1280, 1000), (1000, 1280)):
        for W in order:
            H, n = 800, 2
            game = orc.Game(path, "E1M1", W, H)
            views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::577], n)
            ctx = drr.Context(W, H, 0, n)
            scene = drr.Scene(path, "E1M1", W, H)
            scene.upload_assets(ctx)
            _draw_and_check(ctx, scene, game, views, "%dx%d" % (W, H))
            ctx.close()


def test_recording_the_next_batch_while_the_previous_uploads():
    """drr_submit returns while its copies out of the pinned host lists may still run; resetting and recording the next
    batch right away must not disturb them (the library waits for the copies before it touches the lists)."""
    path, gm = common.wad("e1m1")
    W, H, n = 320, 200, 256
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::7], 2 * n)
    ctx = drr.Context(W, H, 0, 2 * n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    ctx.set_knob("submit_chunks", 8)
    for rep in range(3):
        ctx.reset()
        assert scene.emit_views(ctx, views[:n], first_slot=0) == []
        ctx.submit()                                   # asynchronous: copies + kernels queued
        ctx.reset()                                    # ... and the lists are reused at once
        assert scene.emit_views(ctx, views[n:], first_slot=n) == []
        ctx.submit()
        ctx.sync()
        crcs = ctx.read_checksums(0, 2 * n)  # (a draw zeroes every slot's checksum first: only the second batch's are left)
        for k in range(n, 2 * n, 5):
            v = views[k]
            assert int(crcs[k]) == drr.checksum_numpy(game.render(float(v[0]), float(v[1]), float(v[2]))), (rep, k)
        for k in range(0, n, 17):  # the first batch's frames are still in their slots
            v = views[k]
            _compare(ctx, k, game.render(float(v[0]), float(v[1]), float(v[2])), "first batch, pass %d, view %d" % (rep, k))


# ---- presentation / export (the reference presents Pixels.pixels, src/game.rs:500-533) ----------------------------
@pytest.mark.parametrize("W,H,phases", [(320, 200, 3), (640, 400, 7)])
def test_two_batches_in_flight_on_one_device(W, H, phases):
    """The throughput pattern of INTEGRATION.md 6 and of bench.py's `e2e`: two contexts on ONE device (own stream, own framebuffers),
    each driven by its own host thread through reset -> device front-end -> draw -> checksums, several passes over different
    viewpoint sets, concurrently.  Every checksum must equal the oracle's frame, whatever the other context is doing."""
    import threading
    path, gm = common.wad("e1m1")
    n = 48
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::41], 2 * n)
    want = [drr.checksum_numpy(game.render(float(v[0]), float(v[1]), float(v[2]), phases=phases)) for v in views]
    errors = []

    def work(d):
        try:
            ctx = drr.Context(W, H, 0, n)
            scene = drr.Scene(path, "E1M1", W, H)
            scene.upload_assets(ctx)
            for rep in range(6):
                half = (d + rep) % 2  # the two threads swap viewpoint sets every pass
                ctx.reset()
                assert scene.emit_views_device(ctx, views[half * n:(half + 1) * n], 0.0, phases) == []
                ctx.draw()
                got = ctx.read_checksums(0, n)
                assert [int(x) for x in got] == want[half * n:(half + 1) * n], "thread %d pass %d" % (d, rep)
            ctx.close()
            scene.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]


@pytest.mark.parametrize("W,H", [(320, 200), (324, 203), (1280, 800)])
def test_export_png_and_crc32_of_device_frames(W, H, tmp_path):
    """A device frame through the export path: drr_read_framebuffer -> PNG file -> decoded == the oracle's frame, byte for
    byte; and the device-side CRC-32 of the resident frames == zlib.crc32 of the oracle's frames (324x203: a frame whose size
    is no multiple of the CRC chunk or of 16 bytes)."""
    import zlib
    from doom_rust_renderer_b200 import png
    path, gm = common.wad("e1m1")
    game = orc.Game(path, "E1M1", W, H)
    n = 5
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 4096)[::700], n)
    ctx = drr.Context(W, H, 0, n + 2)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views, first_slot=1) == []  # slots 0 and n+1 stay untouched: all-zero frames
    ctx.submit()
    ctx.sync()
    refs = [game.render(float(v[0]), float(v[1]), float(v[2])) for v in views]
    crcs = ctx.read_crc32(0, n + 2)
    assert int(crcs[0]) == int(crcs[n + 1]) == zlib.crc32(bytes(W * H * 3))
    for k, ref in enumerate(refs):
        assert int(crcs[k + 1]) == zlib.crc32(ref.tobytes()), k
    f = tmp_path / "frame.png"
    f.write_bytes(png.encode_png(ctx.read_framebuffer(3)))
    assert (png.decode_png(f.read_bytes()) == refs[2]).all()
    assert (ctx.read_crc32(2, 1) == crcs[2:3]).all()


def test_real_wad_if_supplied():
    """BASELINE.md:47: with DRR_WAD=<doom1.wad> the E1M1-class workloads use the real IWAD (src/wad.rs:86-109 is the format).
    Skipped when the variable is not set (no WAD ships in the image); with it, frames of a tour over the map's thing positions
    are compared with the oracle like every other workload."""
    from doom_rust_renderer_b200 import workloads
    if not workloads.real_wad():
        pytest.skip("DRR_WAD not set")
    content = workloads.Content("e1m1")
    W, H, n = 320, 200, 24
    game = orc.Game(content.path, "E1M1", W, H)
    views = common.usable_views(game, content.viewpoints(4 * n), n)
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(content.path, "E1M1", W, H)
    scene.upload_assets(ctx)
    _draw_and_check(ctx, scene, game, views, "real WAD")


# ---- BASELINE configs [2], [3], [4] at (or near) their stated sizes ----------------------------------------------------------------
def _oracle_checksums(path, W, H, phases, views, procs=8):
    """Per-frame checksums of the oracle's frames of `views`, several oracle processes side by side."""
    import bench  # the CPU pool of bench.py (oracle worker processes)
    pool = bench.CpuPool(path, W, H, min(procs, len(views)))
    try:
        return pool.render(np.asarray(views, np.float32), phases, want_sums=True)[1]
    finally:
        pool.close()


def _device_batch(kind, W, H, n, phases):
    """n viewpoints of the workload through the device front-end (nudging what the reference would panic on, like bench.py)."""
    import bench
    from doom_rust_renderer_b200 import workloads
    content = workloads.Content(kind, cache_dir=common.CACHE)
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(content.path, "E1M1", W, H)
    scene.upload_assets(ctx)
    used = bench.settle_views_device(scene, ctx, content.viewpoints(n), phases)
    ctx.draw()
    ctx.sync()
    return content.path, ctx, scene, used


@pytest.mark.parametrize("phases,what", [(1, "walls only"), (2, "flats and sky only")])
def test_config2_phase_isolation_at_1280x800(phases, what):
    """BASELINE configs[2]: the column / span isolation at the resolution it is quoted on; 256 viewpoints, every 4th frame
    against the oracle (64 frames, checksums), six of them byte for byte."""
    W, H, n = 1280, 800, 256
    path, ctx, scene, used = _device_batch("e1m1", W, H, n, phases)
    crcs = ctx.read_checksums(0, n)
    idx = np.arange(0, n, 4)
    want = _oracle_checksums(path, W, H, phases, used[idx])
    assert [int(crcs[k]) for k in idx] == want, what
    game = orc.Game(path, "E1M1", W, H)
    for k in idx[::11]:
        v = used[k]
        _compare(ctx, int(k), game.render(float(v[0]), float(v[1]), float(v[2]), phases=phases), "%s, view %d" % (what, k))


def test_config3_things_and_masked_at_4096_viewpoints():
    """BASELINE configs[3] at its full size: 640x400, all phases (things, masked mid-textures, sector light), 4096 viewpoints;
    96 seeded frames against the oracle by checksum, and drawing the batch again from host-recorded lists gives the same 4096
    checksums (device front-end == host front-end at full size)."""
    W, H, n = 640, 400, 4096
    path, ctx, scene, used = _device_batch("e1m1", W, H, n, 7)
    crcs = ctx.read_checksums(0, n)
    idx = np.sort(np.random.default_rng(0xD00D1993).choice(n, 96, replace=False))
    assert [int(crcs[k]) for k in idx] == _oracle_checksums(path, W, H, 7, used[idx])
    ctx.reset()
    assert scene.emit_views(ctx, used, 0.0, 7) == []
    ctx.submit()
    ctx.sync()
    assert (ctx.read_checksums(0, n) == crcs).all()


def test_config4_stress_map_at_1024_viewpoints():
    """BASELINE configs[4]: the stress map at 1920x1200 (three row bands per tile column, hundreds of sprites per view), 1024
    viewpoints of one GPU's share; 64 seeded frames against the oracle by checksum, two byte for byte; a second draw of the
    resident lists changes nothing."""
    W, H, n = 1920, 1200, 1024
    path, ctx, scene, used = _device_batch("stress", W, H, n, 7)
    crcs = ctx.read_checksums(0, n)
    idx = np.sort(np.random.default_rng(0xD00D1993).choice(n, 64, replace=False))
    assert [int(crcs[k]) for k in idx] == _oracle_checksums(path, W, H, 7, used[idx], procs=16)
    game = orc.Game(path, "E1M1", W, H)
    for k in idx[:2]:
        v = used[k]
        _compare(ctx, int(k), game.render(float(v[0]), float(v[1]), float(v[2])), "stress view %d" % k)
    ctx.draw()
    ctx.sync()
    assert (ctx.read_checksums(0, n) == crcs).all()
