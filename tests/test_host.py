"""CPU tests of the product's host side: the C-ABI library loads and exports what include/drr.h declares, the recorded
lists are what was emitted, and the host front-end + column binning reproduce the oracle frame when replayed on the CPU.
No compute call needs a GPU here; drawing without CUDA must fail loudly."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import common
from common import drr, orc, synth_wad


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(common.ROOT, "include", "drr.h")).read()
    declared = set(re.findall(r"\b(drr_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"drr_ctx", "drr_scene"}
    L = ctypes.CDLL(drr.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(drr.EXPORTED_SYMBOLS) <= declared
    assert len(declared) >= 30
    # the product library carries no test infrastructure; the test build is the same library plus the accessors
    import subprocess
    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return {line.split()[-1] for line in out.splitlines() if line.strip()}
    prod, test = exported(drr.LIB_PATH), exported(drr.TEST_LIB_PATH)
    assert not [s for s in prod if "drr_test_" in s or "fastdiv" in s], "test infrastructure in the product library"
    assert any(s.startswith("drr_test_") for s in test) and declared <= test


def test_struct_layouts_match_header():
    assert ctypes.sizeof(drr.DrrView) == 24
    assert ctypes.sizeof(drr.DrrSegHdr) == 48
    assert ctypes.sizeof(drr.DrrVisplaneHdr) == 12
    assert drr.COL_DTYPE.itemsize == 10
    assert drr.DrrSegHdr.start_x.offset == 28 and drr.DrrSegHdr.offset_y.offset == 46


def test_no_cpu_fallback():
    """Without a CUDA device context creation fails with DRR_E_CUDA; the recording-only test context cannot draw."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(drr.DrrError) as e:
        drr.Context(320, 200, 0, 1)
    assert e.value.code == -3
    ctx = drr.Context(64, 64, 0, 1, _host_only=True)
    ctx.upload_palette(np.zeros(768, np.uint8))
    ctx.frame_begin(0, 0, 0, 0, 0, 1, 0)
    ctx.frame_end()
    for fn in (ctx.submit, ctx.draw, ctx.upload_lists, ctx.sync, lambda: ctx.read_framebuffer(0)):
        with pytest.raises(drr.DrrError) as e:
            fn()
        assert e.value.code == -3


def test_product_does_not_import_oracle():
    """The product package must never load, import or execute anything under oracle/."""
    pkg = os.path.join(common.ROOT, "doom_rust_renderer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("oracle-free", ""), os.path.join(root, f)


@pytest.mark.parametrize("W,H,kind,n", [(160, 100, "e1m1", 10), (320, 200, "e1m1", 3), (96, 64, "tiny", 6), (200, 120, "stress", 2)])
def test_front_end_and_binning_replay_equals_oracle(W, H, kind, n):
    """Closure on the CPU: product front-end -> C ABI -> column-binned spans, replayed span by span with the oracle's
    leaf drawers, equals the oracle's direct render of the same viewpoint."""
    path, gm = common.wad(kind)
    game = orc.Game(path, "E1M1", W, H)
    src = synth_wad.walk_viewpoints(gm, 512) if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, 64)
    views = common.usable_views(game, src[:: max(1, len(src) // (n + 2))], n)
    ctx = drr.Context(W, H, 0, len(views), _host_only=True)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]))
        got = common.replay_binned_frame(ctx, k)
        assert (got == ref).all(), "view %d: %d pixels differ" % (k, int((got != ref).any(2).sum()))


def test_front_end_emits_the_oracles_leaf_calls():
    """The product front-end must make the same leaf calls, in the same order, with the same arguments as the oracle's
    restatement of the reference (bitmap handles differ, so they are compared through their texel contents)."""
    path, gm = common.wad("e1m1")
    W, H = 320, 200
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 512)[3::60], 5)
    scene = drr.Scene(path, "E1M1", W, H)
    for v in views:
        game.render(float(v[0]), float(v[1]), float(v[2]), trace=True)
        trace = game.trace()
        # oracle trace through the ABI
        ctx_a = drr.Context(W, H, 0, 1, _host_only=True)
        common.upload_oracle_assets(ctx_a, game)
        common.emit_trace(ctx_a, game, 0, float(v[0]), float(v[1]), float(v[2]), trace)
        # product front-end through the ABI
        ctx_b = drr.Context(W, H, 0, 1, _host_only=True)
        scene.upload_assets(ctx_b)
        scene.emit_view(ctx_b, 0, float(v[0]), float(v[1]), float(v[2]))
        sa, sb = ctx_a.stats(), ctx_b.stats()
        for key in ("seg_headers", "column_records", "visplanes", "visplane_columns", "spans"):
            assert sa[key] == sb[key], (key, sa[key], sb[key])
        assets_a, assets_b = common.AssetsFromCtx(ctx_a), common.AssetsFromCtx(ctx_b)
        segs_a, segs_b = ctx_a._list(1, drr.SEG_DTYPE), ctx_b._list(1, drr.SEG_DTYPE)
        for ga, gb in zip(segs_a, segs_b):
            for f in drr.SEG_DTYPE.names[1:]:
                if f == "tex_base":  # position in the context's texel pool: depends on the upload order
                    continue
                assert ga[f].tobytes() == gb[f].tobytes(), f
            assert (assets_a.bitmap(int(ga["bitmap_slot"])) == assets_b.bitmap(int(gb["bitmap_slot"]))).all()
        assert ctx_a._list(7, drr.COL_DTYPE).tobytes() == ctx_b._list(7, drr.COL_DTYPE).tobytes()  # every drr_col, in order
        assert ctx_a._list(8, np.uint32).tobytes() == ctx_b._list(8, np.uint32).tobytes()          # every (top, bottom) pair
        pa, pb = ctx_a._list(2, drr.PLANE_DTYPE), ctx_b._list(2, drr.PLANE_DTYPE)
        for qa, qb in zip(pa, pb):
            for f in ("height", "light_level", "left", "right"):
                assert qa[f] == qb[f]
            assert (qa["flat_slot"] < 0) == (qb["flat_slot"] < 0)
            if qa["flat_slot"] >= 0:
                assert (assets_a.flat(int(qa["flat_slot"])) == assets_b.flat(int(qb["flat_slot"]))).all()
        spa, spb = ctx_a._list(3, drr.SPAN_DTYPE), ctx_b._list(3, drr.SPAN_DTYPE)
        assert spa.tobytes() == spb.tobytes()
        assert ctx_a._list(0, drr.VIEW_DTYPE).tobytes() == ctx_b._list(0, drr.VIEW_DTYPE).tobytes()


def test_recording_state_machine_and_validation():
    ctx = drr.Context(64, 32, 0, 2, _host_only=True)
    ctx.upload_palette(np.zeros(768, np.uint8))
    ctx.upload_bitmap(1, np.zeros((4, 4), np.int16))
    with pytest.raises(drr.DrrError):
        ctx.upload_bitmap(1, np.zeros((4, 4), np.int16))  # duplicate id
    with pytest.raises(drr.DrrError):
        ctx.upload_bitmap(2, np.full((2, 2), 300, np.int16))  # texel out of range
    with pytest.raises(drr.DrrError):
        ctx.set_sky(1)  # not 256x128
    with pytest.raises(drr.DrrError):
        ctx.emit_visplane(drr.DrrVisplaneHdr(0, 0, 0, 0, 0, 0), np.zeros(1, np.int16), np.zeros(1, np.int16))  # outside a frame
    ctx.frame_begin(0, 0, 0, 0, 0, 1, 0)
    with pytest.raises(drr.DrrError):
        ctx.frame_begin(1, 0, 0, 0, 0, 1, 0)  # nested
    with pytest.raises(drr.DrrError):
        ctx.emit_visplane(drr.DrrVisplaneHdr(5, 0, 0, 0, 0, 0), np.zeros(1, np.int16), np.zeros(1, np.int16))  # unknown flat
    with pytest.raises(drr.DrrError):
        ctx.emit_visplane(drr.DrrVisplaneHdr(-1, 0, 0, 0, 0, 0), np.zeros(1, np.int16), np.zeros(1, np.int16))  # sky not set
    # columns outside the screen are dropped like Pixels::set drops them; ragged / empty inputs are fine
    cols = np.array([(-3, 0, 5, 5, 0), (64, 0, 5, 5, 0), (5, 10, 3, 3, 10), (6, -4, 40, 40, -4)], dtype=drr.COL_DTYPE)
    ctx.emit_columns(drr.DrrSegHdr(1, 100, 0, 1, 1, 2, -1, 0, 0, 10, -8, 8, 0, 0), cols)
    ctx.emit_columns(drr.DrrSegHdr(1, 100, 0, 1, 1, 2, -1, 0, 0, 10, -8, 8, 0, 0), np.zeros(0, drr.COL_DTYPE))
    ctx.frame_end()
    with pytest.raises(drr.DrrError):
        ctx.frame_begin(0, 0, 0, 0, 0, 1, 0)  # slot already recorded
    spans = ctx._list(3, drr.SPAN_DTYPE)
    assert len(spans) == 1 and spans[0]["x"] == 6 and spans[0]["y0"] == 0 and spans[0]["y1"] == 31
    st = ctx.stats()
    assert st["drawlist_bytes_algorithmic"] == 24 + 2 * 48 + 4 * 10
    ctx.reset()
    assert ctx.stats()["frames"] == 0


def test_sass_keeps_products_and_sums_apart():
    """ptxas (CUDA 12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both are explicitly rounded, which would
    change results (one rounding instead of two).  drr_tile.cu adds such products with fma(p, one, c), `one` coming from
    the kernel arguments; here the shipped SASS is checked for it: the tile kernels use the packed instructions, every
    `one` multiply is there (an FFMA2 whose multiplier is a uniform register), and no truncating add got fused."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", drr.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels = re.split(r"\n\s*Function : ", sass)
    tile = [k for k in kernels if "drr_tile_kernel" in k.split("\n", 1)[0]]
    assert len(tile) == 6  # 3 register budgets (4 / 5 / 6 CTAs per SM) x 2 write-out paths
    for k in tile:
        name = k.split("\n", 1)[0]
        assert "FMUL2" in k and "FFMA2" in k and "FADD2" in k, name
        # (c*f) + 2^23 must stay FMUL then FADD.RZ (the only FFMA.RZ allowed is the one inside the IEEE division subroutine)
        assert "FFMA2.RZ" not in k and not re.search(r"FFMA\.RZ [^;]*8388608", k), name
        assert len(re.findall(r"FADD2?\.RZ [^;]*8388608", k)) >= 4, name
        # the `one` multiplies: wall sum, rx, ry, factor (the multiplier is a register or a uniform register, never an immediate 1)
        assert not re.search(r"FFMA2 [^;]*, 1, ", k), name
        assert "UBLKCP" in k and "SYNCS" in k, name                       # palette image: bulk copy on an mbarrier
        assert ("UTMASTG" in k) == ("ELb1E" in name), name               # TMA write-out in the fast-store kernels only


def test_front_end_ptx_has_no_fused_multiply_add():
    """drr_frontend.cuh is plain C++ shared with the host compiler: its f32 expressions are only bit-equal to the host front-end's
    (and the reference's) when nvcc does not contract a*b+c.  Compile the translation unit to PTX with the Makefile's flags and
    look: no fma.rn.f32 of the compiler's making (IEEE division and square root stay div.rn / sqrt.rn in PTX)."""
    import shutil, tempfile
    if not shutil.which("nvcc"):
        pytest.skip("no nvcc")
    src = os.path.join(common.ROOT, "doom_rust_renderer_b200", "csrc")
    flags = re.search(r"^NVFLAGS\s*=\s*(.*)$", open(os.path.join(src, "Makefile")).read(), re.M).group(1)
    assert "-fmad=false" in flags and "-prec-div=true" in flags and "-ftz=false" in flags
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "fe.ptx")
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
                        "-ftz=false", "-ptx", os.path.join(src, "drr_frontend.cu"), "-o", out], check=True, capture_output=True)
        ptx = open(out).read()
    assert "drr_frontend_kernel" in ptx and "div.rn.f32" in ptx
    # the only fma allowed is inside CUDA's fmodf (exact by definition: the remainder is computed with fma on purpose), which
    # mo_pre calls twice (map_objects.rs:44-58): 8 per call, in the two instantiations of Frame::mo_pre
    func, where = "?", {}
    for line in ptx.splitlines():
        m = re.match(r"\.(?:visible |weak )?(?:entry|func)\s.*?(_Z\w+)", line)
        if m:
            func = m.group(1)
        if re.search(r"\bfma\.", line):
            where[func] = where.get(func, 0) + 1
    assert where and all("mo_pre" in f and n == 16 for f, n in where.items()), where


def test_checksum_definitions_agree():
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 192000, 1000 * 3):
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert drr.checksum_host(b) == drr.checksum_numpy(b)


def test_checksum_known_answer():
    """Hand-derived value of the checksum of include/drr.h.  Frame = 52 bytes = the little-endian words 1, 2, ..., 13: two groups
    of 12 words, the second zero-padded.  With C = 0x9E3779B1 = 2654435761:
      s_0 = sum_{j=0..11} (j + 1) (2 j + 1) C = (2 * 506 + 3 * 66 + 12) C = 1222 C = 3243720499942 = 755 * 2^32 + 1020191462
      s_1 = 13 * 1 * C = 34507664893 = 8 * 2^32 + 147926525
      weights of the groups: 1 * C = 2654435761,  2 * C mod 2^32 = 5308871522 - 2^32 = 1013904226
      checksum = 1020191462 * 2654435761 + 147926525 * 1013904226 mod 2^64 = 2858016028634667232"""
    b = np.arange(1, 14, dtype="<u4").view(np.uint8)
    assert b.size == 52
    assert drr.checksum_numpy(b) == 2858016028634667232
    assert drr.checksum_host(b) == 2858016028634667232
    assert drr.checksum_numpy(np.zeros(48, np.uint8)) == 0
    # a change of any single word changes the sum (every multiplier is odd)
    for i in range(13):
        c = b.copy()
        c[4 * i] ^= 1
        assert drr.checksum_numpy(c) != 2858016028634667232


def test_threaded_front_end_records_the_same_lists():
    """drr_scene_emit_views (worker threads, one recorder each, appended in view order) records byte for byte what calling
    drr_scene_emit_view view by view records; a recorder refuses what a context refuses; appending twice the same view fails."""
    path, gm = common.wad("e1m1")
    W, H, n = 160, 100, 70
    views = np.array(synth_wad.walk_viewpoints(gm, 512)[::7][:n], np.float32)
    scene = drr.Scene(path, "E1M1", W, H)
    ctx_seq = drr.Context(W, H, 0, n, _host_only=True)
    scene.upload_assets(ctx_seq)
    skipped_seq = []
    for k, v in enumerate(views):
        try:
            scene.emit_view(ctx_seq, k, float(v[0]), float(v[1]), float(v[2]))
        except drr.DrrError as e:
            assert e.code == -7
            skipped_seq.append(k)
    for threads in (1, 3, 8):
        ctx_mt = drr.Context(W, H, 0, n, _host_only=True)
        scene.upload_assets(ctx_mt)
        assert scene.emit_views(ctx_mt, views, threads=threads) == skipped_seq
        for which, dt in ((0, drr.VIEW_DTYPE), (1, drr.SEG_DTYPE), (2, drr.PLANE_DTYPE), (5, np.uint32), (6, np.uint32), (7, drr.COL_DTYPE),
                          (8, np.uint32), (9, np.uint32), (10, np.uint32)):
            assert ctx_mt._list(which, dt).tobytes() == ctx_seq._list(which, dt).tobytes(), (threads, which)
        assert ctx_mt.stats() == ctx_seq.stats()
        with pytest.raises(drr.DrrError):  # every view index is taken now
            scene.emit_views(ctx_mt, views[:9], threads=2)
        assert ctx_mt.stats() == ctx_seq.stats()  # and the failed append left nothing behind
    # the raw recorder API: same validation as the context's
    L = drr._lib()
    rec = ctypes.c_void_p()
    assert L.drr_recorder_create(ctx_seq.h, ctypes.byref(rec)) == 0
    v = drr.DrrView(0, 0, 0, 0, 1, 0)
    assert L.drr_recorder_emit_columns(rec, ctypes.byref(drr.DrrSegHdr()), None, 0) == -2     # outside a frame
    assert L.drr_recorder_frame_begin(rec, n, ctypes.byref(v)) == -1                            # view index out of range
    assert L.drr_recorder_frame_begin(rec, 0, ctypes.byref(v)) == 0
    assert L.drr_recorder_frame_begin(rec, 1, ctypes.byref(v)) == -2                            # frame not ended
    hdr = drr.DrrSegHdr(10 ** 6, 0, 0, 1, 1, 2, -1, 0, 0, 10, -8, 8, 0, 0)
    assert L.drr_recorder_emit_columns(rec, ctypes.byref(hdr), None, 0) == -5                   # unknown bitmap
    assert b"unknown bitmap" in L.drr_recorder_last_error(rec)
    assert L.drr_append(ctx_seq.h, rec) == -2                                                   # recorder still inside a frame
    assert L.drr_recorder_frame_end(rec) == 0
    assert L.drr_append(ctx_seq.h, rec) == (-1 if 0 not in skipped_seq else 0)                  # view 0 is already recorded
    L.drr_recorder_destroy(rec)


def test_png_export_roundtrip():
    """Presentation/export: an RGB24 frame written as PNG decodes to the same bytes."""
    from doom_rust_renderer_b200 import png
    path, gm = common.wad("tiny")
    game = orc.Game(path, "E1M1", 96, 64)
    x, y, a = game.player_start()
    img = game.render(x, y, a)
    data = png.encode_png(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    assert (png.decode_png(data) == img).all()
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    assert (png.decode_png(png.encode_png(noise)) == noise).all()
    with pytest.raises(ValueError):
        png.encode_png(np.zeros((4, 4), np.uint8))


def test_workload_content_and_real_wad_switch(tmp_path, monkeypatch):
    """workloads.Content: synthetic by default; DRR_WAD points the E1M1-class workloads at a real IWAD, whose THINGS lump
    (read with the module's own directory walk) gives the viewpoint tour.  The synthetic WAD stands in for the real one here."""
    from doom_rust_renderer_b200 import workloads
    monkeypatch.delenv("DRR_WAD", raising=False)
    c = workloads.Content("e1m1", cache_dir=str(tmp_path))
    assert c.real is None and "synthetic" in c.source and len(c.viewpoints(7)) == 7
    things = workloads.wad_things(c.path)
    assert things.shape[1] == 5 and len(things) >= 100 and (things[:, 3] == 1).sum() == 1  # one Player1Start
    monkeypatch.setenv("DRR_WAD", c.path)
    r = workloads.Content("e1m1")
    assert r.real == c.path and r.path == c.path and "real IWAD" in r.source
    v = r.viewpoints(2 * len(things) + 3)
    assert v.shape == (2 * len(things) + 3, 3) and (v[:len(things), :2] == things[:, :2]).all()
    assert not np.allclose(v[0, 2], v[len(things), 2])  # the second round looks elsewhere
    assert workloads.Content("stress").real is None     # only the E1M1-class workloads switch
    monkeypatch.setenv("DRR_WAD", str(tmp_path / "missing.wad"))
    with pytest.raises(FileNotFoundError):
        workloads.Content("e1m1")
