"""ctypes wrapper of tests/host_emu/libvremu.so: the device core compiled for the host.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

EMU_DIR = Path(__file__).resolve().parent / "host_emu"
_lib = None
AUX_DTYPE = np.dtype([
    ("hit", "<i4", (3,)), ("face", "u1"), ("status", "u1"), ("flags", "u1"), ("hit_type", "u1"),
    ("steps_first", "<u4"), ("steps_total", "<u4"), ("node_fetches", "<u4"), ("lookups", "<u4"),
])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", str(EMU_DIR)], check=True, capture_output=True)
        _lib = C.CDLL(str(EMU_DIR / "libvremu.so"))
        _lib.emu_raycast.restype = C.c_int
    return _lib


def set_collapse(on: bool) -> None:
    """solid-subtree collapse of every 64-tree the emulation builds (vr_native_collapse_solid): on by default, as in the library"""
    lib().emu_set_collapse(C.c_int(1 if on else 0))


def raycast(scene, ray_table: np.ndarray, bias=(0, 0, 0), use_svo: bool = False, max_distance: int | None = None,
            shadow_lights: int = 1):
    w, h = scene.width, scene.height
    vol = np.ascontiguousarray(scene.volume, dtype=np.int8)
    lights = np.ascontiguousarray(scene.lights, dtype=np.float32)
    atlas = np.ascontiguousarray(scene.atlas, dtype=np.uint8)
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    rgba[...] = (255, 255, 255, 100)
    aux = np.zeros((h, w), dtype=AUX_DTYPE)
    b = (C.c_int32 * 3)(*[int(v) for v in bias])
    fp = C.POINTER(C.c_float)
    rc = lib().emu_raycast(
        C.c_int(w), C.c_int(h), ray_table.ctypes.data_as(fp), vol.ctypes.data_as(C.c_void_p), C.c_int(scene.n),
        scene.cam_pos.ctypes.data_as(fp), scene.cam_dir.ctypes.data_as(fp), b, lights.ctypes.data_as(fp),
        atlas.ctypes.data_as(C.c_void_p), C.c_int(atlas.shape[1]), C.c_int(atlas.shape[0]), C.c_int(scene.tile), C.c_int(scene.tile),
        C.c_int(scene.max_distance if max_distance is None else max_distance), C.c_int(int(use_svo)), C.c_int(shadow_lights),
        rgba.ctypes.data_as(C.c_void_p), aux.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"emu_raycast failed: {rc}")
    return rgba, aux


def _tree(fn, *args):
    cap_nodes, cap_types = 1 << 22, 1 << 26
    nodes = np.zeros((cap_nodes, 4), dtype=np.uint32)
    types = np.zeros(cap_types, dtype=np.uint8)
    ntypes, levels = C.c_long(0), C.c_int(0)
    fn.restype = C.c_long
    n = fn(*args, nodes.ctypes.data_as(C.c_void_p), C.c_long(cap_nodes), types.ctypes.data_as(C.c_void_p), C.c_long(cap_types),
           C.byref(ntypes), C.byref(levels))
    if n < 0:
        raise RuntimeError(f"tree build failed: {n}")
    return nodes[:n].copy(), types[: ntypes.value].copy(), levels.value


def tree_from_dense(volume: np.ndarray):
    vol = np.ascontiguousarray(volume, dtype=np.int8)
    return _tree(lib().emu_tree_from_dense, vol.ctypes.data_as(C.c_void_p), C.c_int(vol.shape[0]))


def tree_from_columns(lo: np.ndarray, hi: np.ndarray, voxel_type: int = 5):
    lo = np.ascontiguousarray(lo, dtype=np.int32)
    hi = np.ascontiguousarray(hi, dtype=np.int32)
    return _tree(lib().emu_tree_from_columns, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), C.c_int(lo.shape[0]), C.c_int(voxel_type))


def add_chain(t: np.ndarray, d: np.ndarray, n: np.ndarray):
    t = np.ascontiguousarray(t, np.float32); d = np.ascontiguousarray(d, np.float32); n = np.ascontiguousarray(n, np.int32)
    a = np.empty_like(t); b = np.empty_like(t)
    fp = C.POINTER(C.c_float)
    lib().emu_add_chain(t.ctypes.data_as(fp), d.ctypes.data_as(fp), n.ctypes.data_as(C.c_void_p), C.c_int(t.size), a.ctypes.data_as(fp), b.ctypes.data_as(fp))
    return a, b


def tree_from_ref(desc: np.ndarray, root: int, dim: int):
    d = np.ascontiguousarray(desc, dtype=np.uint64)
    return _tree(lib().emu_tree_from_ref, d.ctypes.data_as(C.c_void_p), C.c_uint64(d.size), C.c_uint64(root), C.c_int(dim))


def grid_from_tree(nodes: np.ndarray, levels: int, dim: int, directed: bool = False):
    """host version of the closed-form walk's top grid (vr_octree.cpp): vr_native_grid -> (uint32[G,G,G], shift, bits);
    directed: vr_native_grid_directed -> (uint32[8,G,G,G], shift, bits), one table per direction octant"""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    cap = 1 << (27 if directed else 24)
    out = np.zeros(cap, dtype=np.uint32)
    gs, gb = C.c_int(0), C.c_int(0)
    fn = lib().emu_grid_directed_from_tree if directed else lib().emu_grid_from_tree
    fn.restype = C.c_long
    n = fn(nodes.ctypes.data_as(C.c_void_p), C.c_int(levels), C.c_int(dim), out.ctypes.data_as(C.c_void_p), C.c_long(cap), C.byref(gs), C.byref(gb))
    if n < 0:
        raise RuntimeError(f"grid build failed: {n}")
    g = 1 << gb.value
    return (out[:n].reshape(8, g, g, g) if directed else out[:n].reshape(g, g, g)).copy(), gs.value, gb.value
