"""ctypes binding of the CPU oracle (oracle/drr_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (doom_rust_renderer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PHASE_WALLS, PHASE_PLANES, PHASE_MASKED, PHASE_ALL = 1, 2, 4, 7


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "drr_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_game_new.restype = C.c_void_p
        L.orc_game_new.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.orc_game_free.argtypes = [C.c_void_p]
        L.orc_player_start.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_floor_height_at.restype = C.c_float
        L.orc_floor_height_at.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_sector_at.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_render.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int]
        L.orc_set_tic.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.orc_sector_lights.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_object_states.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_palette.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_bitmap_count.argtypes = [C.c_void_p]
        L.orc_bitmap_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_bitmap_texels.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sky_bitmap_id.argtypes = [C.c_void_p]
        L.orc_flat_count.argtypes = [C.c_void_p]
        L.orc_flat_texels.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_trace_count.argtypes = [C.c_void_p]
        L.orc_trace_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_trace_visplane_arrays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_diminish_color.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_leaf_new.restype = C.c_void_p
        L.orc_leaf_new.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_leaf_free.argtypes = [C.c_void_p]
        L.orc_leaf_column.argtypes = [C.c_void_p] * 5
        L.orc_leaf_visplane.argtypes = [C.c_void_p] * 8
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    """The reference would have panicked on this input."""


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Game:
    """Headless equivalent of the reference's Game::new + Renderer (src/game.rs:118-196, src/renderer/mod.rs)."""

    def __init__(self, wad_path: str, map_name: str, W: int, H: int):
        self.L = lib()
        self.W, self.H = W, H
        self.h = self.L.orc_game_new(wad_path.encode(), map_name.encode(), W, H)
        if not self.h:
            raise OracleError(self.L.orc_last_error().decode())

    def close(self):
        if self.h:
            self.L.orc_game_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def player_start(self):
        out = np.zeros(3, np.float32)
        if self.L.orc_player_start(self.h, _p(out)) != 0:
            raise OracleError(self.L.orc_last_error().decode())
        return float(out[0]), float(out[1]), float(out[2])

    def floor_height_at(self, x, y) -> float:
        return float(self.L.orc_floor_height_at(self.h, x, y))

    def sector_at(self, x, y) -> int:
        return int(self.L.orc_sector_at(self.h, x, y))

    def set_tic(self, tic: int, seed: int = 0):
        """The world `tic` game ticks after the start (thinkers.rs / lights.rs / map_objects.rs:63-95; PCG32 stream `seed`)."""
        if self.L.orc_set_tic(self.h, tic, seed) != 0:
            raise OracleError(self.L.orc_last_error().decode())

    def world_state(self):
        lights = np.zeros(1 << 16, np.int16)
        n = self.L.orc_sector_lights(self.h, _p(lights))
        objs = np.zeros((1 << 16, 4), np.int32)
        m = self.L.orc_object_states(self.h, _p(objs))
        return lights[:n].copy(), objs[:m].copy()

    def render(self, x, y, angle, timestamp=0.0, phases=PHASE_ALL, trace=False, out=None) -> np.ndarray:
        if out is None:
            out = np.empty((self.H, self.W, 3), np.uint8)
        rc = self.L.orc_render(self.h, x, y, angle, timestamp, _p(out), phases, 1 if trace else 0)
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())
        return out

    # ---- assets ----
    def palette(self) -> np.ndarray:
        out = np.empty(768, np.uint8)
        self.L.orc_palette(self.h, _p(out))
        return out

    def bitmap_count(self) -> int:
        return self.L.orc_bitmap_count(self.h)

    def bitmap(self, i) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        assert self.L.orc_bitmap_size(self.h, i, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.int16)
        self.L.orc_bitmap_texels(self.h, i, _p(out))
        return out

    def sky_bitmap_id(self) -> int:
        return self.L.orc_sky_bitmap_id(self.h)

    def flat_count(self) -> int:
        return self.L.orc_flat_count(self.h)

    def flat(self, i) -> np.ndarray:
        out = np.empty(4096, np.uint8)
        assert self.L.orc_flat_texels(self.h, i, _p(out)) == 0
        return out

    # ---- trace of the leaf calls made by the last render(trace=True) ----
    def trace(self):
        n = self.L.orc_trace_count(self.h)
        calls = []
        ints = np.zeros(17, np.int32)
        flts = np.zeros(7, np.float32)
        for i in range(n):
            self.L.orc_trace_get(self.h, i, _p(ints), _p(flts))
            d = dict(kind=int(ints[0]), phase=int(ints[1]), asset=int(ints[2]), light_level=int(ints[3]), start_x=int(ints[4]),
                     end_x=int(ints[5]), offset_x=int(ints[6]), offset_y=int(ints[7]), x=int(ints[8]), clipped_bottom_y=int(ints[9]),
                     clipped_top_y=int(ints[10]), bottom_y=int(ints[11]), top_y=int(ints[12]), is_sky=int(ints[13]), height=int(ints[14]),
                     left=int(ints[15]), right=int(ints[16]), line=flts[:4].copy(), start_offset=np.float32(flts[4]),
                     bottom_height=np.float32(flts[5]), top_height=np.float32(flts[6]))
            if d["kind"] == 1:
                top = np.empty(self.W, np.int16)
                bottom = np.empty(self.W, np.int16)
                self.L.orc_trace_visplane_arrays(self.h, i, _p(top), _p(bottom))
                d["top"], d["bottom"] = top, bottom
            calls.append(d)
        return calls


class Leaf:
    """The three leaf drawers on caller-supplied arguments (bitmap_render.rs:213-276, visplanes.rs:42-130)."""

    def __init__(self, W: int, H: int, palette768: np.ndarray):
        self.L = lib()
        self.W, self.H = W, H
        pal = np.ascontiguousarray(palette768, np.uint8)
        self.h = self.L.orc_leaf_new(W, H, _p(pal))

    def __del__(self):
        try:
            if self.h:
                self.L.orc_leaf_free(self.h)
                self.h = None
        except Exception:
            pass

    def column(self, pixels, texels, light_level, line, start_offset, start_x, end_x, bottom_height, top_height, offset_x, offset_y, x,
               clipped_bottom_y, clipped_top_y, bottom_y, top_y):
        texels = np.ascontiguousarray(texels, np.int16)
        h, w = texels.shape
        ints = np.array([w, h, light_level, start_x, end_x, offset_x, offset_y, x, clipped_bottom_y, clipped_top_y, bottom_y, top_y], np.int32)
        flts = np.array([line[0], line[1], line[2], line[3], start_offset, bottom_height, top_height], np.float32)
        rc = self.L.orc_leaf_column(self.h, _p(pixels), _p(texels), _p(ints), _p(flts))
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())

    def visplane(self, pixels, flat, sky_texels, top, bottom, is_sky, height, light_level, left, right, pos_x, pos_y, floor_height, angle):
        ints = np.array([is_sky, height, light_level, left, right], np.int32)
        flts = np.array([pos_x, pos_y, floor_height, angle], np.float32)
        top = np.ascontiguousarray(top, np.int16)
        bottom = np.ascontiguousarray(bottom, np.int16)
        fl = np.ascontiguousarray(flat, np.uint8) if flat is not None else None
        sk = np.ascontiguousarray(sky_texels, np.int16) if sky_texels is not None else None
        rc = self.L.orc_leaf_visplane(self.h, _p(pixels), _p(fl) if fl is not None else None, _p(sk) if sk is not None else None, _p(top),
                                      _p(bottom), _p(ints), _p(flts))
        if rc != 0:
            raise OracleError(self.L.orc_last_error().decode())


def diminish_color(rgb, light_level: int, distance: int):
    i = np.array(rgb, np.uint8)
    o = np.zeros(3, np.uint8)
    lib().orc_diminish_color(_p(i), light_level, distance, _p(o))
    return tuple(int(v) for v in o)
