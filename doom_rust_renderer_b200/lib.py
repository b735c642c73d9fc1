"""ctypes binding of libdrr.so (include/drr.h).

The binding is deliberately thin: every method is one C-ABI call.  There is no Python or CPU implementation of the draw
path; if the shared library is missing this module raises, and without a CUDA device `Context(...)` raises DrrError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrr.so")            # the product
TEST_LIB_PATH = os.path.join(_HERE, "libdrr_test.so")  # the product's objects + the drr_test_* accessors (-DDRR_TESTING)
_USE_TEST_LIBRARY = False


def use_test_library():
    """tests/conftest.py: load libdrr_test.so instead of libdrr.so (must be called before the first call into the library)."""
    global _USE_TEST_LIBRARY
    if _LIB is not None and not _USE_TEST_LIBRARY:
        raise RuntimeError("libdrr.so is already loaded")
    _USE_TEST_LIBRARY = True


PHASES_WALLS, PHASES_PLANES, PHASES_MASKED, PHASES_ALL = 1, 2, 4, 7
PHASE_WALL, PHASE_MASKED = 0, 2
FLAT_SKY = -1


class DrrView(C.Structure):
    _fields_ = [("pos_x", C.c_float), ("pos_y", C.c_float), ("floor_height", C.c_float), ("angle", C.c_float),
                ("cos_angle", C.c_float), ("sin_angle", C.c_float)]


class DrrSegHdr(C.Structure):
    _fields_ = [("bitmap_id", C.c_int32), ("light_level", C.c_int16), ("phase", C.c_int16),
                ("line_start_x", C.c_float), ("line_start_y", C.c_float), ("line_end_x", C.c_float), ("line_end_y", C.c_float),
                ("start_offset", C.c_float), ("start_x", C.c_int32), ("end_x", C.c_int32),
                ("bottom_height", C.c_float), ("top_height", C.c_float), ("offset_x", C.c_int16), ("offset_y", C.c_int16)]


class DrrVisplaneHdr(C.Structure):
    _fields_ = [("flat_id", C.c_int16), ("height", C.c_int16), ("light_level", C.c_int16), ("left", C.c_int16), ("right", C.c_int16),
                ("reserved", C.c_int16)]


class DrrStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("frames", "seg_headers", "column_records", "visplanes", "visplane_columns",
                                          "drawlist_bytes_algorithmic", "device_list_bytes", "spans", "kernel_launches")]


assert C.sizeof(DrrView) == 24 and C.sizeof(DrrSegHdr) == 48 and C.sizeof(DrrVisplaneHdr) == 12

# numpy dtype of drr_col (10 bytes)
COL_DTYPE = np.dtype([("x", "<i2"), ("clipped_top_y", "<i2"), ("clipped_bottom_y", "<i2"), ("bottom_y", "<i2"), ("top_y", "<i2")])
assert COL_DTYPE.itemsize == 10


class DrrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (_lib().drr_error_name(code).decode() if _LIB is not None else code, msg))
        self.code = code


_LIB = None


def _lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = TEST_LIB_PATH if _USE_TEST_LIBRARY else LIB_PATH
        if not os.path.exists(path):
            raise ImportError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or `make -C doom_rust_renderer_b200/csrc`). There is no fallback implementation." % os.path.basename(path))
        L = C.CDLL(path)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        sig = {
            "drr_ctx_create": (i, [i, i, i, i, C.POINTER(vp)]), "drr_ctx_destroy": (None, [vp]),
            "drr_last_error": (C.c_char_p, [vp]), "drr_error_name": (C.c_char_p, [i]),
            "drr_set_stream": (i, [vp, vp]), "drr_get_stream": (vp, [vp]), "drr_set_knob": (i, [vp, C.c_char_p, i]),
            "drr_upload_palette": (i, [vp, vp]), "drr_upload_bitmap": (i, [vp, i, i, i, vp]), "drr_upload_flat": (i, [vp, i, vp]),
            "drr_set_sky": (i, [vp, i]), "drr_reset": (i, [vp]), "drr_frame_begin": (i, [vp, i, C.POINTER(DrrView)]),
            "drr_emit_columns": (i, [vp, C.POINTER(DrrSegHdr), vp, i]),
            "drr_emit_visplane": (i, [vp, C.POINTER(DrrVisplaneHdr), vp, vp]), "drr_frame_end": (i, [vp]), "drr_frame_abort": (i, [vp]),
            "drr_upload_lists": (i, [vp]), "drr_draw": (i, [vp]), "drr_submit": (i, [vp]), "drr_sync": (i, [vp]),
            "drr_read_framebuffer": (i, [vp, i, vp]), "drr_read_checksums": (i, [vp, i, i, vp]), "drr_read_crc32": (i, [vp, i, i, vp]),
            "drr_checksum_host": (C.c_uint64, [vp, C.c_uint64]), "drr_get_stats": (i, [vp, C.POINTER(DrrStats)]),
            "drr_time_draw": (i, [vp, i, C.POINTER(f), C.POINTER(f), C.POINTER(f)]),
            "drr_profile_begin": (i, [vp, i]), "drr_profile_end": (i, [vp, C.POINTER(i), C.POINTER(f), C.POINTER(f)]),
            "drr_scene_load": (i, [C.c_char_p, C.c_char_p, i, i, C.POINTER(vp)]), "drr_scene_free": (None, [vp]),
            "drr_scene_last_error": (C.c_char_p, [vp]), "drr_scene_upload_assets": (i, [vp, vp]),
            "drr_scene_player_start": (i, [vp, vp]), "drr_scene_emit_view": (i, [vp, vp, i, f, f, f, f, i]),
            "drr_scene_emit_views": (i, [vp, vp, i, vp, i, f, i, i, vp]),
            "drr_recorder_create": (i, [vp, C.POINTER(vp)]), "drr_recorder_destroy": (None, [vp]), "drr_recorder_last_error": (C.c_char_p, [vp]),
            "drr_recorder_frame_begin": (i, [vp, i, C.POINTER(DrrView)]), "drr_recorder_emit_columns": (i, [vp, C.POINTER(DrrSegHdr), vp, i]),
            "drr_recorder_emit_visplane": (i, [vp, C.POINTER(DrrVisplaneHdr), vp, vp]), "drr_recorder_frame_end": (i, [vp]),
            "drr_recorder_frame_abort": (i, [vp]), "drr_append": (i, [vp, vp]),
            "drr_fe_upload_map": (i, [vp, vp]), "drr_fe_emit_views": (i, [vp, i, vp, i, i, vp]),
            "drr_fe_last_times": (i, [vp, C.POINTER(f), C.POINTER(f)]), "drr_fe_last_mode": (i, [vp]), "drr_fe_map_id": (C.c_uint32, [vp]),
            "drr_scene_emit_views_device": (i, [vp, vp, i, vp, i, f, i, vp]),
            "drr_scene_set_tic": (i, [vp, C.c_uint32, C.c_uint64]), "drr_scene_counts": (i, [vp, C.POINTER(i), C.POINTER(i)]),
            "drr_scene_world_state": (i, [vp, vp, vp]),
            "drr_test_fe_emit_views_host": (i, [vp, i, vp, i, i, vp]), "drr_test_fe_download_lists": (i, [vp]),
            "drr_test_ctx_create_host_only": (i, [i, i, i, C.POINTER(vp)]),
            "drr_test_list": (vp, [vp, i, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "drr_test_bitmap_info": (i, [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "drr_test_bitmap_texels": (i, [vp, i, vp]), "drr_test_flat_texels": (i, [vp, i, vp]), "drr_test_palette": (i, [vp, vp]),
            "drr_test_sky_slot": (i, [vp]), "drr_test_tile_config": (i, [vp, C.POINTER(i), C.POINTER(i)]),
            "drr_test_bitmap_id_of_slot": (i, [vp, i]), "drr_test_flat_id_of_slot": (i, [vp, i]),
            "drr_test_device_bins": (i, [vp, vp, vp]), "drr_test_tile_bands": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "drr_test_fastdiv": (i, [vp, i, C.c_longlong, C.c_longlong, f, C.c_uint32, C.c_uint32, C.POINTER(C.c_ulonglong), vp]),
        }
        for name, (res, args) in sig.items():
            if name.startswith("drr_test_") and not _USE_TEST_LIBRARY:
                continue  # not in the product library
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


EXPORTED_SYMBOLS = [
    "drr_ctx_create", "drr_ctx_destroy", "drr_last_error", "drr_error_name", "drr_set_stream", "drr_get_stream", "drr_set_knob",
    "drr_upload_palette", "drr_upload_bitmap", "drr_upload_flat", "drr_set_sky", "drr_reset", "drr_frame_begin",
    "drr_emit_columns", "drr_emit_visplane", "drr_frame_end", "drr_frame_abort", "drr_upload_lists", "drr_draw", "drr_submit", "drr_sync",
    "drr_read_framebuffer", "drr_read_checksums", "drr_read_crc32", "drr_checksum_host", "drr_get_stats", "drr_time_draw",
    "drr_profile_begin", "drr_profile_end",
    "drr_scene_load", "drr_scene_free", "drr_scene_last_error", "drr_scene_upload_assets", "drr_scene_player_start",
    "drr_scene_emit_view", "drr_scene_emit_views",
    "drr_recorder_create", "drr_recorder_destroy", "drr_recorder_last_error", "drr_recorder_frame_begin", "drr_recorder_emit_columns",
    "drr_recorder_emit_visplane", "drr_recorder_frame_end", "drr_recorder_frame_abort", "drr_append",
    "drr_fe_upload_map", "drr_fe_emit_views", "drr_fe_last_times", "drr_fe_last_mode", "drr_fe_map_id", "drr_scene_emit_views_device",
    "drr_scene_set_tic", "drr_scene_counts", "drr_scene_world_state",
]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def checksum_host(frame: np.ndarray) -> int:
    a = np.ascontiguousarray(frame, np.uint8)
    return int(_lib().drr_checksum_host(_ptr(a), a.size))


from .checksum import checksum_numpy  # noqa: E402,F401  (pure numpy; lives apart so that CPU-only tools need not import this module)


class Context:
    """drr_ctx: one GPU, `max_views` RGB24 framebuffers resident in HBM."""

    def __init__(self, width: int, height: int, device: int = 0, max_views: int = 1, _host_only: bool = False):
        self.L = _lib()
        self.W, self.H, self.max_views = width, height, max_views
        h = C.c_void_p()
        if _host_only:  # CPU-test recording context: bins draw lists, can never draw
            rc = self.L.drr_test_ctx_create_host_only(width, height, max_views, C.byref(h))
        else:
            rc = self.L.drr_ctx_create(width, height, device, max_views, C.byref(h))
        if rc != 0:
            raise DrrError(rc, self.L.drr_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.drr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise DrrError(rc, self.L.drr_last_error(self.h).decode())

    # ---- assets
    def upload_palette(self, rgb768):
        a = np.ascontiguousarray(rgb768, np.uint8).reshape(-1)
        assert a.size == 768
        self._ck(self.L.drr_upload_palette(self.h, _ptr(a)))

    def upload_bitmap(self, bitmap_id: int, texels):
        a = np.ascontiguousarray(texels, np.int16)
        h, w = a.shape
        self._ck(self.L.drr_upload_bitmap(self.h, bitmap_id, w, h, _ptr(a)))

    def upload_flat(self, flat_id: int, px4096):
        a = np.ascontiguousarray(px4096, np.uint8).reshape(-1)
        assert a.size == 4096
        self._ck(self.L.drr_upload_flat(self.h, flat_id, _ptr(a)))

    def set_sky(self, bitmap_id: int):
        self._ck(self.L.drr_set_sky(self.h, bitmap_id))

    # ---- recording
    def reset(self):
        self._ck(self.L.drr_reset(self.h))

    def frame_begin(self, view_idx: int, pos_x, pos_y, floor_height, angle, cos_angle, sin_angle):
        v = DrrView(pos_x, pos_y, floor_height, angle, cos_angle, sin_angle)
        self._ck(self.L.drr_frame_begin(self.h, view_idx, C.byref(v)))

    def emit_columns(self, hdr: DrrSegHdr, cols: np.ndarray):
        cols = np.ascontiguousarray(cols, COL_DTYPE)
        self._ck(self.L.drr_emit_columns(self.h, C.byref(hdr), _ptr(cols), cols.size))

    def emit_visplane(self, hdr: DrrVisplaneHdr, top, bottom):
        t = np.ascontiguousarray(top, np.int16)
        b = np.ascontiguousarray(bottom, np.int16)
        assert t.size == b.size == hdr.right - hdr.left + 1
        self._ck(self.L.drr_emit_visplane(self.h, C.byref(hdr), _ptr(t), _ptr(b)))

    def frame_end(self):
        self._ck(self.L.drr_frame_end(self.h))

    def frame_abort(self):
        self._ck(self.L.drr_frame_abort(self.h))

    # ---- execution
    def set_knob(self, name: str, value: int):
        """Tuning / diagnostic knob (drr.h: drr_set_knob); the DRR_* environment variables are only read at creation."""
        self._ck(self.L.drr_set_knob(self.h, name.encode(), int(value)))

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.drr_set_stream(self.h, C.c_void_p(cuda_stream)))

    def upload_lists(self):
        self._ck(self.L.drr_upload_lists(self.h))

    def draw(self):
        self._ck(self.L.drr_draw(self.h))

    def submit(self):
        self._ck(self.L.drr_submit(self.h))

    def sync(self):
        self._ck(self.L.drr_sync(self.h))

    def read_framebuffer(self, view_idx: int) -> np.ndarray:
        out = np.empty((self.H, self.W, 3), np.uint8)
        self._ck(self.L.drr_read_framebuffer(self.h, view_idx, _ptr(out)))
        return out

    def read_checksums(self, first: int, count: int) -> np.ndarray:
        out = np.zeros(count, np.uint64)
        self._ck(self.L.drr_read_checksums(self.h, first, count, _ptr(out)))
        return out

    def read_crc32(self, first: int, count: int) -> np.ndarray:
        """zlib.crc32 of each resident frame's W*H*3 bytes, computed on the device (export side)."""
        out = np.zeros(count, np.uint32)
        self._ck(self.L.drr_read_crc32(self.h, first, count, _ptr(out)))
        return out

    def stats(self) -> dict:
        s = DrrStats()
        self._ck(self.L.drr_get_stats(self.h, C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in DrrStats._fields_}

    def time_draw(self, iters: int):
        t, s, m = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.drr_time_draw(self.h, iters, C.byref(t), C.byref(s), C.byref(m)))
        return t.value, s.value, m.value

    def kernel_name(self) -> str:
        """The dominant kernel (csrc/drr_tile.cu): 32-column tiles, 8 lanes per span; <resident CTAs per SM the register
        budget is set for, TMA write-out>."""
        return "drr_tile_kernel<4|6, true>"

    def tile_bands(self):
        """(nbands, band_rows, nlists): the row bands of a tile and the number of span lists per column the bin kernel writes."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.drr_test_tile_bands(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def device_bins(self, nframes: int):
        """(colidx [nframes][nlists][W] of COLIDX_DTYPE, recs [n][2] u32 = (y0 | y1 << 16, kind | flags)) as the bin kernel wrote
        them: one span list per (column, row band)."""
        nlists = self.tile_bands()[2]
        ci = np.zeros(nframes * nlists * self.W, COLIDX_DTYPE)
        recs = np.zeros((max(1, int(self._list(5, np.uint32)[-1]) * nlists), 2), np.uint32)  # one slot per emitted column and list
        self._ck(self.L.drr_test_device_bins(self.h, _ptr(ci), _ptr(recs)))
        return ci.reshape(nframes, nlists, self.W), recs

    def profile_begin(self, max_steps: int):
        self._ck(self.L.drr_profile_begin(self.h, max_steps))

    def profile_end(self):
        """(steps, setup_ms_total, march_ms_total) of the draws since profile_begin."""
        n, s, m = C.c_int(), C.c_float(), C.c_float()
        self._ck(self.L.drr_profile_end(self.h, C.byref(n), C.byref(s), C.byref(m)))
        return n.value, s.value, m.value

    def test_fastdiv(self, mode: int, n0: int, n1: int, cfy: float = 0.0, lo: int = 0, stride: int = 1):
        """Device self-check of the hoisted-reciprocal division against __fdiv_rn; returns (mismatches, first offending pair)."""
        bad = C.c_ulonglong()
        first = np.zeros(2, np.float32)
        self._ck(self.L.drr_test_fastdiv(self.h, mode, n0, n1, cfy, lo, stride, C.byref(bad), _ptr(first)))
        return bad.value, (float(first[0]), float(first[1]))

    # ---- device front-end
    def fe_emit_views(self, views: np.ndarray, phases: int = 3, first_slot: int = 0, _on_host: bool = False):
        """drr_fe_emit_views on the map uploaded with drr_fe_upload_map (Scene.emit_views_device does both).  `_on_host` runs
        the same per-view code on the CPU into the host lists (test infrastructure).  Returns the view indices the reference
        would have panicked on."""
        v = np.ascontiguousarray(np.asarray(views, np.float32).reshape(-1, 3))
        status = np.zeros(len(v), np.int32)
        fn = self.L.drr_test_fe_emit_views_host if _on_host else self.L.drr_fe_emit_views
        self._ck(fn(self.h, first_slot, _ptr(v), len(v), phases, _ptr(status)))
        return [first_slot + int(k) for k in np.nonzero(status == -7)[0]]

    def fe_last_times(self):
        """(count_ms, emit_ms) of the last fe_emit_views; in single-pass mode (fe_last_mode() == 1) the first is the gather of the frames' View records (the compaction kernel with DRR_FE_COMPACT=1)."""
        a, b = C.c_float(), C.c_float()
        self._ck(self.L.drr_fe_last_times(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fe_last_mode(self) -> int:
        """1 = single pass (per-view slabs, read in place by the draw kernels), 2 = count pass + emit pass."""
        return int(self.L.drr_fe_last_mode(self.h))

    def fe_download_lists(self):
        """Test infrastructure: copy the device-written lists back so that _list() can show them."""
        self._ck(self.L.drr_test_fe_download_lists(self.h))

    # ---- test-only views of the recorded lists (CPU-testable host logic)
    def _list(self, which: int, dtype) -> np.ndarray:
        n, sz = C.c_uint64(), C.c_uint64()
        p = self.L.drr_test_list(self.h, which, C.byref(n), C.byref(sz))
        dt = np.dtype(dtype)
        assert dt.itemsize == sz.value, (dt.itemsize, sz.value)
        if not p or n.value == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (n.value * sz.value)).from_address(p)
        return np.frombuffer(buf, dt).copy()


VIEW_DTYPE = np.dtype([("pos_x", "<f4"), ("pos_y", "<f4"), ("floor_height", "<f4"), ("angle", "<f4"), ("cos_a", "<f4"), ("sin_a", "<f4")])
SEG_DTYPE = np.dtype([("bitmap_slot", "<u4"), ("light_level", "<i2"), ("phase", "<i2"), ("lsx", "<f4"), ("lsy", "<f4"), ("lex", "<f4"),
                      ("ley", "<f4"), ("start_offset", "<f4"), ("start_x", "<i4"), ("end_x", "<i4"), ("bottom_height", "<f4"),
                      ("top_height", "<f4"), ("offset_x", "<i2"), ("offset_y", "<i2"), ("cols_first", "<u4"), ("n", "<u4"),
                      ("x0", "<i2"), ("x1", "<i2"), ("tex_base", "<u4"), ("tex_w", "<i2"), ("tex_h", "<i2"), ("tex_opaque", "<u4"), ("pad", "<u4", (2,))])
PLANE_DTYPE = np.dtype([("flat_slot", "<i2"), ("height", "<i2"), ("light_level", "<i2"), ("left", "<i2"), ("right", "<i2"), ("kind", "<i2"),
                        ("arr_first", "<u4")])
SPAN_DTYPE = np.dtype([("y0", "<u2"), ("y1", "<u2"), ("x", "<u2"), ("kind", "u1"), ("pad", "u1"), ("op", "<u4"), ("top_y", "<i2"), ("bottom_y", "<i2")])
COLIDX_DTYPE = np.dtype([("first", "<u4"), ("n", "<u4")])
KIND_WALL, KIND_WALL_HOLES, KIND_FLAT, KIND_SKY, KIND_SKY_HOLES = 0, 1, 2, 3, 4


class Scene:
    """drr_scene: the reference's Game::new asset/map loading plus its Renderer front-end, emitting draw lists."""

    def __init__(self, wad_path: str, map_name: str, width: int, height: int):
        self.L = _lib()
        self.W, self.H = width, height
        h = C.c_void_p()
        rc = self.L.drr_scene_load(wad_path.encode(), map_name.encode(), width, height, C.byref(h))
        if rc != 0:
            raise DrrError(rc, self.L.drr_scene_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.drr_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise DrrError(rc, self.L.drr_scene_last_error(self.h).decode())

    def upload_assets(self, ctx: Context):
        self._ck(self.L.drr_scene_upload_assets(self.h, ctx.h))

    def player_start(self):
        out = np.zeros(3, np.float32)
        self._ck(self.L.drr_scene_player_start(self.h, _ptr(out)))
        return float(out[0]), float(out[1]), float(out[2])

    def emit_view(self, ctx: Context, view_idx: int, x: float, y: float, angle: float, timestamp: float = 0.0, phases: int = PHASES_ALL):
        self._ck(self.L.drr_scene_emit_view(self.h, ctx.h, view_idx, x, y, angle, timestamp, phases))

    def emit_views(self, ctx: Context, views: np.ndarray, timestamp: float = 0.0, phases: int = PHASES_ALL, first_slot: int = 0,
                   threads: int = 0):
        """views: float32 [n][3] = (x, y, angle).  The front-end runs on `threads` worker threads (0 = one per host core), each
        recording into its own recorder; frames end up in the context in view order.  Returns the list of view indices the
        reference would have panicked on (nothing is recorded for those; their framebuffer slots keep their previous
        contents)."""
        v = np.ascontiguousarray(np.asarray(views, np.float32).reshape(-1, 3))
        status = np.zeros(len(v), np.int32)
        self._ck(self.L.drr_scene_emit_views(self.h, ctx.h, first_slot, _ptr(v), len(v), timestamp, phases, threads, _ptr(status)))
        return [first_slot + int(k) for k in np.nonzero(status == -7)[0]]  # DRR_E_PANIC

    def emit_views_device(self, ctx: Context, views: np.ndarray, timestamp: float = 0.0, phases: int = PHASES_ALL, first_slot: int = 0,
                          _on_host: bool = False):
        """The same batch with the front-end running ON THE GPU (drr_fe_upload_map + drr_fe_emit_views), every phase.  The
        context must be reset first; afterwards ctx.draw() renders.  `_on_host` (test infrastructure) runs the same per-view
        code on the CPU into the context's host lists."""
        v = np.ascontiguousarray(np.asarray(views, np.float32).reshape(-1, 3))
        status = np.zeros(len(v), np.int32)
        if _on_host:
            self._ck(self.L.drr_scene_emit_views_device(self.h, ctx.h, first_slot, _ptr(v), 0, timestamp, phases, None))  # map only
            return ctx.fe_emit_views(v, phases=phases, first_slot=first_slot, _on_host=True)
        self._ck(self.L.drr_scene_emit_views_device(self.h, ctx.h, first_slot, _ptr(v), len(v), timestamp, phases, _ptr(status)))
        return [first_slot + int(k) for k in np.nonzero(status == -7)[0]]

    def set_tic(self, tic: int, seed: int = 0):
        """The world `tic` game ticks after the start (sector light effects, map-object animation; random draws from a PCG32
        stream seeded with `seed`).  tic 0 = the WAD as loaded."""
        self._ck(self.L.drr_scene_set_tic(self.h, tic, seed))

    def world_state(self):
        """(sector light levels [n_sectors] int16, object states [n_objects][4] int32 = sprite, frame, full_bright, is S_NULL)."""
        ns, no = C.c_int(), C.c_int()
        self._ck(self.L.drr_scene_counts(self.h, C.byref(ns), C.byref(no)))
        lights, objs = np.zeros(ns.value, np.int16), np.zeros((no.value, 4), np.int32)
        self._ck(self.L.drr_scene_world_state(self.h, _ptr(lights), _ptr(objs)))
        return lights, objs

    def upload_map_for_device_front_end(self, ctx: Context, timestamp: float = 0.0):
        """drr_fe_upload_map only (a zero-view drr_scene_emit_views_device)."""
        v = np.zeros((0, 3), np.float32)
        self._ck(self.L.drr_scene_emit_views_device(self.h, ctx.h, 0, _ptr(v), 0, timestamp, PHASES_ALL, None))
