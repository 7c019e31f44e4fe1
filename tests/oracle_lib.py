"""ctypes wrapper of the CPU parity oracle (oracle/libvroracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
_lib = None

ST_SKIP_PRIMARY, ST_OOB, ST_MAXDIST, ST_SHADOW_HIT, ST_SKIP_REDIRECT, ST_BOUNCES = range(6)
FL_LIT, FL_REFLECTED, FL_TIE, FL_ATLAS_CLAMP, FL_FRAC0, FL_NEAR, FL_NEAR_AIR = 1, 2, 4, 8, 16, 32, 64

AUX_DTYPE = np.dtype([
    ("hit", "<i4", (3,)), ("face", "u1"), ("status", "u1"), ("flags", "u1"), ("hit_type", "u1"),
    ("steps_first", "<u4"), ("steps_total", "<u4"), ("pad", "<u4", (2,)),
])
assert AUX_DTYPE.itemsize == 32


class VroScene(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("ray_table", C.POINTER(C.c_float)),
        ("map", C.POINTER(C.c_int8)), ("map_dim", C.c_int32 * 3),
        ("cam_dir", C.c_float * 2), ("cam_pos", C.c_float * 3), ("trig", C.c_float * 4),
        ("lights", C.POINTER(C.c_float)), ("light_count", C.c_int32),
        ("atlas", C.POINTER(C.c_uint8)), ("atlas_dim", C.c_int32 * 2), ("tile_dim", C.c_int32 * 2),
        ("oct_desc", C.POINTER(C.c_uint64)), ("oct_desc_len", C.c_uint64),
        ("octdim", C.c_int64), ("oct_root_index", C.c_int64),
        ("max_distance", C.c_int32), ("shadow_lights", C.c_int32),
        ("col_lo", C.POINTER(C.c_int32)), ("col_hi", C.POINTER(C.c_int32)),
        ("canonical_t", C.c_int32), ("max_bounces", C.c_int32),
        ("tree64", C.POINTER(C.c_uint32)), ("tree64_levels", C.c_int32),
    ]


class VroCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "pixels", "pixels_written", "primary_rays", "shadow_rays", "reflect_rays", "dda_steps",
        "texel_fetches", "svo_desc_fetches", "svo_cell_changes", "tie_pixels")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build() -> Path:
    subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return ORACLE_DIR / "libvroracle.so"


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = ORACLE_DIR / "libvroracle.so"
        if not path.exists():
            build()
        L = C.CDLL(str(path))
        L.vro_make_ray_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.vro_make_ray_table.restype = None
        L.vro_raycast.argtypes = [C.POINTER(VroScene), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_void_p,
                                  C.POINTER(VroCounters), C.c_int, C.c_int]
        L.vro_raycast.restype = C.c_int
        L.vro_get_oct_vox.argtypes = [C.POINTER(C.c_uint64), C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.vro_get_oct_vox.restype = None
        L.vro_octree_generate.argtypes = [C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64)]
        L.vro_octree_generate.restype = C.c_int64
        L.vro_num_procs.restype = C.c_int
        _lib = L
    return _lib


def make_ray_table(width: int, height: int) -> np.ndarray:
    out = np.empty((height, width, 4), dtype=np.float32)
    lib().vro_make_ray_table(width, height, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def octree_generate(volume: np.ndarray, buffer_size: int = 100000) -> tuple[np.ndarray, int, int]:
    """Restated Octree::Generate: (buffer, root_index, descriptors used)."""
    vol = np.ascontiguousarray(volume, dtype=np.int8)
    buf = np.zeros(buffer_size, dtype=np.uint64)
    used = C.c_uint64(0)
    root = lib().vro_octree_generate(vol.ctypes.data_as(C.POINTER(C.c_int8)), vol.shape[0],
                                     buf.ctypes.data_as(C.POINTER(C.c_uint64)), buffer_size, C.byref(used))
    return buf, int(root), int(used.value)


def get_oct_vox(desc: np.ndarray, root: int, dim: int, pos) -> tuple[int, tuple[int, int, int], int, int]:
    p = (C.c_int32 * 3)(*[int(v) for v in pos])
    sub = (C.c_int32 * 3)()
    found, res, scale = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    lib().vro_get_oct_vox(desc.ctypes.data_as(C.POINTER(C.c_uint64)), root, dim, p, C.byref(found), sub, C.byref(res), C.byref(scale))
    return int(found.value), (sub[0], sub[1], sub[2]), int(res.value), int(scale.value)


def trig_of(cam_dir) -> np.ndarray:
    """sinf/cosf of the camera angles with the C library's float routines (same ones the caster calls)."""
    libm = C.CDLL("libm.so.6")
    libm.sinf.restype = libm.cosf.restype = C.c_float
    libm.sinf.argtypes = libm.cosf.argtypes = [C.c_float]
    return np.array([libm.sinf(float(cam_dir[0])), libm.cosf(float(cam_dir[0])),
                     libm.sinf(float(cam_dir[1])), libm.cosf(float(cam_dir[1]))], dtype=np.float32)


def raycast(scene, ray_table: np.ndarray | None = None, octree: tuple[np.ndarray, int] | None = None,
            rows: tuple[int, int] | None = None, want_aux: bool = True, want_counters: bool = False,
            count_svo: bool = False, threads: int = 0, max_distance: int | None = None, row_stride: int = 1,
            shadow_lights: int = 1, canonical_t: bool = False, keep_near: bool = False, max_bounces: int = 0, tree64=None):
    """Runs the restated reference kernel (dense branch) on a scene.Scene.
    Returns (rgba [H,W,4] prefilled with (255,255,255,100), aux or None, counters dict or None)."""
    w, h = scene.width, scene.height
    if ray_table is None:
        ray_table = make_ray_table(w, h)
    columns = getattr(scene, "columns", None)
    vol = np.ascontiguousarray(scene.volume, dtype=np.int8) if scene.volume is not None else None
    lights = np.ascontiguousarray(scene.lights, dtype=np.float32)
    atlas = np.ascontiguousarray(scene.atlas, dtype=np.uint8)
    s = VroScene()
    s.width, s.height = w, h
    s.ray_table = ray_table.ctypes.data_as(C.POINTER(C.c_float))
    if vol is not None:
        s.map = vol.ctypes.data_as(C.POINTER(C.c_int8))
        s.map_dim[:] = [vol.shape[2], vol.shape[1], vol.shape[0]]
    else:
        col_lo = np.ascontiguousarray(columns[0], dtype=np.int32)
        col_hi = np.ascontiguousarray(columns[1], dtype=np.int32)
        s.col_lo = col_lo.ctypes.data_as(C.POINTER(C.c_int32))
        s.col_hi = col_hi.ctypes.data_as(C.POINTER(C.c_int32))
        s.map_dim[:] = [scene.n, scene.n, scene.n]
    s.cam_dir[:] = [float(v) for v in scene.cam_dir]
    s.cam_pos[:] = [float(v) for v in scene.cam_pos]
    s.trig[:] = [float(v) for v in trig_of(scene.cam_dir)]
    s.lights = lights.ctypes.data_as(C.POINTER(C.c_float))
    s.light_count = lights.shape[0]
    s.atlas = atlas.ctypes.data_as(C.POINTER(C.c_uint8))
    s.atlas_dim[:] = [atlas.shape[1], atlas.shape[0]]
    s.tile_dim[:] = [scene.tile, scene.tile]
    if octree is not None:
        desc, root = octree
        s.oct_desc = desc.ctypes.data_as(C.POINTER(C.c_uint64))
        s.oct_desc_len = desc.size
        s.oct_root_index = root
    if tree64 is not None:                   # (nodes uint32[n, 4], levels): descriptor-fetch counters without a reference-format buffer
        t64 = np.ascontiguousarray(tree64[0], dtype=np.uint32)
        s.tree64 = t64.ctypes.data_as(C.POINTER(C.c_uint32))
        s.tree64_levels = int(tree64[1])
    s.octdim = scene.n
    s.max_distance = scene.max_distance if max_distance is None else max_distance
    s.shadow_lights = shadow_lights          # 1 = the reference (light 0 only); > 1 = the multi-light extension
    s.max_bounces = max_bounces              # 0 = the reference's 2 (kernel:357)
    s.canonical_t = 1 if canonical_t else 0  # False = the reference; True = "Oracle-B" (closed-form crossing times)
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    rgba[...] = (255, 255, 255, 100)
    aux = np.zeros((h, w), dtype=AUX_DTYPE) if want_aux else None
    counters = VroCounters() if (want_counters or count_svo) else None
    y0, y1 = rows if rows else (0, h)
    rc = lib().vro_raycast(C.byref(s), y0, y1, row_stride, rgba.ctypes.data_as(C.POINTER(C.c_uint8)),
                           aux.ctypes.data_as(C.c_void_p) if aux is not None else None,
                           C.byref(counters) if counters is not None else None, int(count_svo), threads)
    if rc != 0:
        raise RuntimeError(f"vro_raycast failed: {rc}")
    if aux is not None and not keep_near:
        aux["flags"] &= np.uint8(~(FL_NEAR | FL_NEAR_AIR) & 0xFF)   # the near-tie classifier is only of interest to the walk = 2 tests
    return rgba, aux, (counters.as_dict() if counters is not None else None)


def num_procs() -> int:
    return int(lib().vro_num_procs())
