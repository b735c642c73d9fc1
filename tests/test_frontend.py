"""The device front-end (BSP walk, seg clipping, occlusion arrays and visplane building per viewpoint
on the GPU, csrc/drr_frontend.cuh) against the host front-end (csrc/host/drr_scene.cpp) and against the oracle.

CPU part: the per-viewpoint code is one source compiled for both sides; here it runs on the CPU (drr_test_fe_emit_views_host,
test infrastructure) and must produce, byte for byte, the lists the host front-end records -- views, ops, seg headers, column
records, visplane headers, (top, bottom) rows, per-frame tables and statistics -- including which viewpoints the reference
would panic on.  GPU part (-m gpu): the kernel's lists, downloaded, equal the host front-end's bytes, and the frames drawn
from them equal the oracle's, bit for bit.
"""
from __future__ import annotations

import os
import struct

import numpy as np
import pytest

import common
from common import drr, orc, synth_wad

LISTS = ((0, drr.VIEW_DTYPE), (1, drr.SEG_DTYPE), (2, drr.PLANE_DTYPE), (5, np.uint32), (6, np.uint32), (7, drr.COL_DTYPE), (8, np.uint32),
         (9, np.uint32), (10, np.uint32))
NAMES = {0: "views", 1: "seg headers", 2: "visplane headers", 5: "frame_rec_base", 6: "frame_slot", 7: "column records", 8: "visplane rows",
         9: "ops", 10: "frame_op_base"}


def _assert_same_lists(a: drr.Context, b: drr.Context, what: str):
    for which, dt in LISTS:
        x, y = a._list(which, dt), b._list(which, dt)
        assert x.shape == y.shape, (what, NAMES[which], x.shape, y.shape)
        if x.tobytes() != y.tobytes():
            k = int(np.nonzero(x != y)[0][0])
            raise AssertionError("%s: %s differ first at %d: %s vs %s" % (what, NAMES[which], k, x[k], y[k]))


def _views(kind, gm, n):
    return np.array(synth_wad.walk_viewpoints(gm, n) if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, n), np.float32)


def _wad_without_things(kind: str) -> str:
    """The synthetic IWAD with only the player start left in THINGS (the device front-end draws no map objects, so it accepts
    DRR_PHASES_MASKED only for such a map)."""
    path, _ = common.wad(kind)
    out = os.path.join(common.CACHE, "synth_%s_nothings.wad" % kind)
    data = open(path, "rb").read()
    n, diro = struct.unpack("<II", data[4:12])
    lumps = []
    for i in range(n):
        off, size = struct.unpack("<II", data[diro + 16 * i: diro + 16 * i + 8])
        name = data[diro + 16 * i + 8: diro + 16 * i + 16]
        body = data[off: off + size] if size else b""
        if name.rstrip(b"\0") == b"THINGS":
            body = b"".join(body[k: k + 10] for k in range(0, len(body), 10) if struct.unpack("<h", body[k + 6: k + 8])[0] == 1)
        lumps.append((name, body))
    blob, directory, off = bytearray(), bytearray(), 12
    for name, body in lumps:
        directory += struct.pack("<II", off if body else 0, len(body)) + name
        blob += body
        off += len(body)
    new = b"IWAD" + struct.pack("<II", len(lumps), 12 + len(blob)) + bytes(blob) + bytes(directory)
    if not os.path.exists(out) or open(out, "rb").read() != new:
        with open(out + ".tmp", "wb") as f:
            f.write(new)
        os.replace(out + ".tmp", out)
    return out


# ---- CPU: the shared per-viewpoint code against the host front-end ---------------------------------------------------------
@pytest.mark.parametrize("kind,W,H,n,phases,ts", [("e1m1", 320, 200, 300, 3, 0.0), ("e1m1", 160, 100, 96, 1, 0.0), ("e1m1", 324, 200, 96, 2, 0.4),
                                                  ("e1m1", 1280, 800, 40, 3, 0.7), ("stress", 200, 120, 160, 3, 0.0), ("stress", 640, 400, 24, 3, 0.0)])
def test_front_end_code_emits_the_host_front_ends_lists(kind, W, H, n, phases, ts):
    path, gm = common.wad(kind)
    views = _views(kind, gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, timestamp=ts, phases=phases, threads=2)
    b = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(b)
    scene.upload_map_for_device_front_end(b, ts)
    assert b.fe_emit_views(views, phases=phases, _on_host=True) == skipped
    _assert_same_lists(a, b, "%s %dx%d phases %d" % (kind, W, H, phases))
    assert a.stats() == b.stats()


@pytest.mark.parametrize("kind,W,H,n,things", [("e1m1", 320, 200, 260, True), ("e1m1", 320, 200, 120, False), ("e1m1", 640, 400, 60, True),
                                                ("stress", 200, 120, 120, True), ("e1m1", 1280, 800, 24, True), ("e1m1", 1000, 900, 8, True),
                                                ("e1m1", 96, 2000, 6, True), ("e1m1", 33, 64, 40, True), ("e1m1", 4097, 48, 3, True)])
def test_front_end_code_all_phases(kind, W, H, n, things):
    """Phases C and D: map objects (rotation, projection, clip arrays from the parts in front, depth order, interleave with the
    masked mid-textures behind each sprite) and the remaining masked mid-textures, last created first -- with and without
    things in the map."""
    path, gm = common.wad(kind)
    views = _views(kind, gm, n)
    scene = drr.Scene(path if things else _wad_without_things(kind), "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, phases=7, threads=2)
    b = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, phases=7, _on_host=True) == skipped
    _assert_same_lists(a, b, "%s %dx%d all phases, things=%s" % (kind, W, H, things))
    assert a.stats() == b.stats()
    segs = a._list(1, drr.SEG_DTYPE)
    assert (segs["phase"] == 2).any(), "the batch shows nothing masked: the test would prove nothing"
    if things:  # sprites have no texture offsets and their own bitmaps: the batch must contain some
        bare = drr.Scene(_wad_without_things(kind), "E1M1", W, H)
        c = drr.Context(W, H, 0, n, _host_only=True)
        bare.upload_assets(c)
        bare.emit_views(c, views, phases=7, threads=2)
        assert a.stats()["seg_headers"] > c.stats()["seg_headers"]


PANIC_VIEWS = {  # found by random search (tools: a seg exactly through the eye point makes the reference panic)
    "e1m1": [(-59.0, -320.0, 0.785398), (-1443.0, 896.0, 0.785398)],
    "stress": [(-883.0, 0.0, 0.785398), (1652.0, 768.0, 0.785398), (-1895.0, -512.0, 0.785398)],
}


def _views_with_panics(kind, gm, n):
    v = _views(kind, gm, n)
    bad = np.array(PANIC_VIEWS[kind], np.float32)
    at = np.linspace(3, n - 2, len(bad)).astype(int)
    v[at] = bad
    return v, [int(k) for k in at]


def test_front_end_code_panics_where_the_host_front_end_panics():
    """Viewpoints on which the reference panics (a wall exactly through the eye) get no frame on either path, and the frames
    after them land where the host front-end puts them."""
    W, H, n = 160, 100, 120
    for kind in ("e1m1", "stress"):
        path, gm = common.wad(kind)
        views, at = _views_with_panics(kind, gm, n)
        scene = drr.Scene(path, "E1M1", W, H)
        a = drr.Context(W, H, 0, n, _host_only=True)
        scene.upload_assets(a)
        skipped = scene.emit_views(a, views, phases=3, threads=4)
        assert skipped == at, "the known panicking viewpoints no longer panic: the test would prove nothing"
        b = drr.Context(W, H, 0, n, _host_only=True)
        scene.upload_assets(b)
        scene.upload_map_for_device_front_end(b)
        assert b.fe_emit_views(views, phases=3, _on_host=True) == skipped
        _assert_same_lists(a, b, kind)
        assert a.stats() == b.stats()


def test_front_end_single_pass_two_pass_and_slab_overflow_agree(monkeypatch):
    """Single-pass mode (per-view slabs + compaction), two-pass mode (count pass, exact offsets, emit pass) and the fallback
    from the first to the second when a view outgrows its slab all write the same bytes."""
    W, H, n = 320, 200, 150
    bare = _wad_without_things("e1m1")
    _, gm = common.wad("e1m1")
    views = _views("e1m1", gm, n)
    scene = drr.Scene(bare, "E1M1", W, H)
    ctxs = []
    # (the last case starts with working arrays far too small for the masked phase: they are enlarged and the batch redone)
    for env, mode in (({}, 1), ({"DRR_FE_TWO_PASS": "1"}, 2), ({"DRR_FE_SLAB_DIV": "64"}, 2), ({"DRR_FE_CAP_RENDERS": "16", "DRR_FE_CAP_DSEGS": "2"}, 1)):
        for k in ("DRR_FE_TWO_PASS", "DRR_FE_SLAB_DIV", "DRR_FE_CAP_RENDERS", "DRR_FE_CAP_DSEGS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = drr.Context(W, H, 0, n, _host_only=True)
        scene.upload_assets(c)
        scene.upload_map_for_device_front_end(c)
        assert c.fe_emit_views(views, phases=7, _on_host=True) == []
        assert c.fe_last_mode() == mode, env
        ctxs.append(c)
    _assert_same_lists(ctxs[0], ctxs[1], "single-pass vs two-pass")
    _assert_same_lists(ctxs[0], ctxs[2], "single-pass vs overflow fallback")
    _assert_same_lists(ctxs[0], ctxs[3], "single-pass vs enlarged working arrays")
    assert ctxs[0].stats() == ctxs[1].stats() == ctxs[2].stats() == ctxs[3].stats()


def test_front_end_state_machine():
    W, H, n = 160, 100, 8
    path, gm = common.wad("e1m1")
    views = _views("e1m1", gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    ctx = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(ctx)
    with pytest.raises(drr.DrrError) as e:  # no map yet
        ctx.fe_emit_views(views, _on_host=True)
    assert e.value.code == -2
    scene.upload_map_for_device_front_end(ctx)
    with pytest.raises(drr.DrrError) as e:  # the device path itself needs a GPU: no CPU fallback
        ctx.fe_emit_views(views)
    assert e.value.code == -3
    assert ctx.fe_emit_views(views, _on_host=True) == []
    with pytest.raises(drr.DrrError) as e:  # a batch is already recorded
        ctx.fe_emit_views(views, _on_host=True)
    assert e.value.code == -2
    ctx.reset()
    with pytest.raises(drr.DrrError):  # more views than framebuffers
        ctx.fe_emit_views(np.concatenate([views, views]), _on_host=True)
    assert ctx.fe_emit_views(views[:3], first_slot=5, _on_host=True) == []
    assert list(ctx._list(6, np.uint32)) == [5, 6, 7]


# ---- GPU: the kernel -----------------------------------------------------------------------------------------------------
def _compare(ctx, k, ref, what):
    got = ctx.read_framebuffer(k)
    diff = (got != ref).any(2)
    if diff.any():
        ys, xs = np.nonzero(diff)
        raise AssertionError("%s: %d pixels differ; first at x=%d y=%d got=%s want=%s" % (what, int(diff.sum()), xs[0], ys[0], got[ys[0], xs[0]], ref[ys[0], xs[0]]))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,W,H,n,phases,ts", [("e1m1", 320, 200, 700, 3, 0.0), ("e1m1", 1280, 800, 40, 3, 0.4), ("stress", 200, 120, 200, 3, 0.0),
                                                  ("e1m1", 324, 200, 33, 1, 0.0), ("stress", 1920, 1200, 6, 2, 0.0), ("e1m1", 320, 200, 500, 7, 0.0),
                                                  ("e1m1", 640, 400, 64, 7, 0.7), ("stress", 640, 400, 48, 7, 0.0), ("e1m1", 1280, 800, 16, 4, 0.0),
                                                  ("e1m1", 1000, 900, 6, 7, 0.0), ("e1m1", 96, 2000, 5, 7, 0.0), ("e1m1", 33, 64, 64, 7, 0.0),
                                                  ("e1m1", 4097, 48, 3, 7, 0.0)])
def test_device_front_end_lists_equal_host_front_end(kind, W, H, n, phases, ts):
    path, gm = common.wad(kind)
    views = _views(kind, gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, timestamp=ts, phases=phases)
    b = drr.Context(W, H, 0, n)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, timestamp=ts, phases=phases) == skipped
    b.fe_download_lists()
    _assert_same_lists(a, b, "%s %dx%d phases %d" % (kind, W, H, phases))
    assert a.stats() == {**b.stats(), "kernel_launches": 0, "device_list_bytes": a.stats()["device_list_bytes"]}
    count_ms, emit_ms = b.fe_last_times()
    assert count_ms > 0 and emit_ms > 0
    # (the stress map's sprite-heavy views outgrow the per-view slabs: that batch falls back to two passes)
    assert b.fe_last_mode() == (2 if os.environ.get("DRR_FE_TWO_PASS") or (kind == "stress" and phases & 4) else 1)


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"DRR_FE_TWO_PASS": "1"}, {"DRR_FE_SLAB_DIV": "64"}])
def test_device_front_end_two_pass_and_overflow_fallback(env, monkeypatch):
    W, H, n = 320, 200, 300
    path, gm = common.wad("e1m1")
    views = _views("e1m1", gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, phases=3)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    b = drr.Context(W, H, 0, n)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, phases=3) == skipped
    assert b.fe_last_mode() == 2
    b.fe_download_lists()
    _assert_same_lists(a, b, str(env))


@pytest.mark.gpu
@pytest.mark.parametrize("compact", [False, True])
def test_device_front_end_slabs_drawn_in_place_or_compacted_first(compact, monkeypatch):
    """Single-pass mode: by default the draw kernels read the front-end's per-view slabs in place, DRR_FE_COMPACT=1 copies them into
    dense lists first.  Either way the frames are the same (checksums equal to the host lists' frames), drawing twice changes
    nothing, and the dense lists -- made on demand in the first case, AFTER the draw -- are byte for byte the host front-end's."""
    W, H, n = 320, 200, 90
    path, gm = common.wad("e1m1")
    views = _views("e1m1", gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, phases=7)
    ref = drr.Context(W, H, 0, n)
    scene.upload_assets(ref)
    assert scene.emit_views(ref, views, phases=7) == skipped
    ref.submit()
    want = ref.read_checksums(0, n)
    if compact:
        monkeypatch.setenv("DRR_FE_COMPACT", "1")
    b = drr.Context(W, H, 0, n)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, phases=7) == skipped
    assert b.fe_last_mode() == 1
    b.draw()
    got = b.read_checksums(0, n)
    b.draw()
    assert (b.read_checksums(0, n) == got).all()
    keep = [k for k in range(n) if k not in skipped]
    assert (got[keep] == want[keep]).all()
    b.fe_download_lists()
    _assert_same_lists(a, b, "compact=%s" % compact)
    b.draw()  # and the slabs are still what the draw kernels read
    assert (b.read_checksums(0, n) == got).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,W,H,n,phases", [("e1m1", 320, 200, 24, 3), ("e1m1", 1280, 800, 4, 3), ("stress", 640, 400, 6, 3), ("e1m1", 320, 200, 32, 7),
                                               ("e1m1", 640, 400, 8, 7), ("stress", 1920, 1200, 2, 7),
                                               # widths around the edges of the front-end's shared-memory modes (all arrays / occlusion arrays only / none)
                                               ("e1m1", 896, 504, 6, 7), ("e1m1", 928, 520, 6, 3), ("e1m1", 1600, 400, 4, 7)])
def test_device_front_end_frames_match_oracle(kind, W, H, n, phases):
    """viewpoints -> device front-end -> bin kernel -> tile kernel == the oracle's frames (no list ever touches the host)."""
    path, gm = common.wad(kind)
    game = orc.Game(path, "E1M1", W, H)
    src = synth_wad.walk_viewpoints(gm, 4096)[:: 4096 // (n + 4)] if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, 64)
    views = common.usable_views(game, src, n)
    assert len(views) == n
    ctx = drr.Context(W, H, 0, n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views_device(ctx, views, phases=phases) == []
    ctx.submit()  # nothing to upload: the same as draw()
    ctx.sync()
    crcs = ctx.read_checksums(0, n)
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]), phases=phases)
        _compare(ctx, k, ref, "%s %dx%d view %d" % (kind, W, H, k))
        assert int(crcs[k]) == drr.checksum_numpy(ref)
    # a second batch through the same context, other slots first
    ctx.reset()
    assert scene.emit_views_device(ctx, views[::-1], phases=phases) == []
    ctx.draw()
    ctx.sync()
    assert list(ctx.read_checksums(0, n)) == list(crcs[::-1])


@pytest.mark.gpu
def test_device_front_end_skips_panicking_viewpoints_all_phases():
    W, H, n = 160, 100, 1500
    path, gm = common.wad("stress")
    views, at = _views_with_panics("stress", gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, phases=7)
    assert skipped == at
    b = drr.Context(W, H, 0, n)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, phases=7) == skipped
    b.fe_download_lists()
    _assert_same_lists(a, b, "stress, all phases")
    # and the frames: the host front-end's lists through drr_submit in a second context
    c = drr.Context(W, H, 0, n)
    scene.upload_assets(c)
    assert scene.emit_views(c, views, phases=7) == skipped
    c.submit()
    c.sync()
    b.draw()
    b.sync()
    assert b.read_checksums(0, n).tobytes() == c.read_checksums(0, n).tobytes()
    assert len(set(b.read_checksums(0, n).tolist())) > n // 2


def test_front_end_map_validation():
    """drr_fe_upload_map checks every index the kernel will follow (so the walk needs no bounds tests) and refuses maps it cannot
    walk; errors come back as codes, never as a crash."""
    import ctypes as C
    L = drr._lib()

    class Node(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("dx", C.c_float), ("dy", C.c_float), ("right", C.c_int32), ("left", C.c_int32)]

    class Sub(C.Structure):
        _fields_ = [("first", C.c_int32), ("count", C.c_int32)]

    class Seg(C.Structure):
        _fields_ = [("v1x", C.c_float), ("v1y", C.c_float), ("v2x", C.c_float), ("v2y", C.c_float), ("line", C.c_int32), ("dir", C.c_int16), ("off", C.c_int16)]

    class Line(C.Structure):
        _fields_ = [("front", C.c_int32), ("back", C.c_int32), ("flags", C.c_int32)]

    class Side(C.Structure):
        _fields_ = [("xo", C.c_float), ("yo", C.c_float), ("upper", C.c_int32), ("lower", C.c_int32), ("middle", C.c_int32), ("sector", C.c_int32)]

    class Sector(C.Structure):
        _fields_ = [("f", C.c_int16), ("c", C.c_int16), ("l", C.c_int16), ("sky", C.c_int16), ("ff", C.c_int16), ("cf", C.c_int16), ("fs", C.c_int16), ("cs", C.c_int16)]

    class Thing(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("angle", C.c_float), ("sector", C.c_int32), ("fb", C.c_int32), ("rotate", C.c_int32),
                    ("bitmap", C.c_int32 * 8), ("top", C.c_int16 * 8)]

    class Map(C.Structure):
        _fields_ = [("nodes", C.POINTER(Node)), ("n_nodes", C.c_int32), ("subs", C.POINTER(Sub)), ("n_subs", C.c_int32), ("segs", C.POINTER(Seg)), ("n_segs", C.c_int32),
                    ("lines", C.POINTER(Line)), ("n_lines", C.c_int32), ("sides", C.POINTER(Side)), ("n_sides", C.c_int32),
                    ("sectors", C.POINTER(Sector)), ("n_sectors", C.c_int32), ("things", C.POINTER(Thing)), ("n_things", C.c_int32)]

    assert (C.sizeof(Node), C.sizeof(Sub), C.sizeof(Seg), C.sizeof(Line), C.sizeof(Side), C.sizeof(Sector), C.sizeof(Thing)) == (24, 8, 24, 12, 24, 16, 72)
    ctx = drr.Context(64, 40, 0, 4, _host_only=True)

    def good():
        m = Map()
        m.nodes, m.n_nodes = (Node * 1)(Node(0, 0, 1, 0, -1, -2)), 1   # children: subsectors 0 and 1
        m.subs, m.n_subs = (Sub * 2)(Sub(0, 1), Sub(1, 1)), 2
        m.segs, m.n_segs = (Seg * 2)(Seg(64, -32, 64, 32, 0, 0, 0), Seg(64, 32, 64, -32, 0, 1, 0)), 2
        m.lines, m.n_lines = (Line * 1)(Line(0, -1, 1)), 1
        m.sides, m.n_sides = (Side * 1)(Side(0, 0, -1, -1, -1, 0)), 1
        m.sectors, m.n_sectors = (Sector * 1)(Sector(0, 128, 160, 0, -2, -2, 0, 0)), 1
        return m

    m = good()
    assert L.drr_fe_upload_map(ctx.h, C.byref(m)) == 0
    v = np.array([[0, 0, 0]], np.float32)
    # a well-formed two-seg map walks: either an (empty) frame or -- a facing wall in a sector whose flat lumps are missing
    # makes the reference panic in Flats::get -- no frame; never an error
    assert ctx.fe_emit_views(v, phases=7, _on_host=True) in ([], [0])
    for breakit, code in ((lambda m: setattr(m.nodes[0], "left", 5), -1), (lambda m: setattr(m.nodes[0], "left", 0), -1), (lambda m: setattr(m.subs[1], "count", 2), -1), (lambda m: setattr(m.segs[0], "line", 1), -1),
                          (lambda m: setattr(m.lines[0], "front", 3), -1), (lambda m: setattr(m.sides[0], "sector", -1), -1), (lambda m: setattr(m, "n_segs", 0), -1),
                          (lambda m: setattr(m, "n_things", 1), -1)):
        m = good()
        breakit(m)
        assert L.drr_fe_upload_map(ctx.h, C.byref(m)) == code
        ctx.reset()
        with pytest.raises(drr.DrrError):  # a refused map leaves no map behind
            ctx.fe_emit_views(v, _on_host=True)
    m = good()
    t = Thing(0, 0, 0, 0, 0, 0)
    t.bitmap[0] = 7  # never uploaded
    m.things, m.n_things = (Thing * 1)(t), 1
    assert L.drr_fe_upload_map(ctx.h, C.byref(m)) == -5


@pytest.mark.gpu
def test_device_front_end_configs1_full_size():
    """BASELINE configs[1] at full size (4096 viewpoints of the walk, 320x200, walls + flats + sky): the device front-end's
    frames equal the host front-end's frame for frame (checksums), a seeded sample of them equals the oracle's pixels, and a
    second run of the batch is idempotent."""
    W, H, n = 320, 200, 4096
    path, gm = common.wad("e1m1")
    views = _views("e1m1", gm, n)
    scene = drr.Scene(path, "E1M1", W, H)
    a = drr.Context(W, H, 0, n)
    scene.upload_assets(a)
    skipped = scene.emit_views(a, views, phases=3)
    a.submit()
    want = a.read_checksums(0, n)
    a.reset()
    assert scene.emit_views_device(a, views, phases=3) == skipped
    assert a.fe_last_mode() == 1
    a.draw()
    got = a.read_checksums(0, n)
    ok = np.ones(n, bool)
    ok[[k for k in skipped]] = False
    assert (got[ok] == want[ok]).all()
    assert len(set(got[ok].tolist())) > n // 2
    game = orc.Game(path, "E1M1", W, H)
    for k in np.random.default_rng(0xD00D1993).choice(np.nonzero(ok)[0], 12, replace=False):
        ref = game.render(float(views[k][0]), float(views[k][1]), float(views[k][2]), phases=3)
        assert int(got[k]) == drr.checksum_numpy(ref), k
        _compare(a, int(k), ref, "view %d" % k)
    a.reset()
    assert scene.emit_views_device(a, views, phases=3) == skipped
    a.draw()
    assert (a.read_checksums(0, n) == got).all()
