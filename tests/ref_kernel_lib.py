"""ctypes wrapper of oracle/_ref/libref_kernel*.so: the REFERENCE's own kernels/ray_caster_kernel.cl compiled for the
CPU by g++ through oracle/ref_shim/cl_shim.h (recipe: `make -C oracle ref`, run by __graft_entry__.build() wherever
/root/reference exists; the built libraries travel to the GPU box, the reference sources do not).
TEST INFRASTRUCTURE ONLY: it pins the oracle (tests/test_reference_kernel.py) and nothing else uses it."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import oracle_lib as O

REF_DIR = Path(__file__).resolve().parent.parent / "oracle" / "_ref"
_libs: dict[str, C.CDLL] = {}


def _name(lifted) -> str:
    """lifted: False = the kernel verbatim; True = max_distance lifted; "wide" = max_distance lifted and the three
    private stacks widened from 8 to 32 entries (maps deeper than 256^3)"""
    return {False: "libref_kernel.so", True: "libref_kernel_md.so", "wide": "libref_kernel_wide.so"}[lifted]


def available(lifted=False) -> bool:
    return (REF_DIR / _name(lifted)).exists()


def lib(lifted) -> C.CDLL:
    name = _name(lifted)
    if name not in _libs:
        _libs[name] = C.CDLL(str(REF_DIR / name))
        _libs[name].ref_raycast.restype = C.c_int
        _libs[name].ref_kernel_max_distance.restype = C.c_int
    return _libs[name]


def raycast(scene, ray_table: np.ndarray | None = None, octree: tuple[np.ndarray, int] | None = None, lifted=False,
            octenabled: int = 1, row_stride: int = 1, rows: tuple[int, int] | None = None):
    """One frame through the reference kernel.  lifted=False: the kernel verbatim (max_distance 20, kernel:326);
    lifted=True: max_distance = scene.max_distance.  Returns (rgba prefilled with (255,255,255,100) like host:280-286,
    written mask).  `octree` = (descriptors, root index) in the reference layout; without it a one-entry buffer with an
    all-empty root is bound (get_oct_vox then reports the whole map as one empty cell, bias as the kernel computes it)."""
    w, h = scene.width, scene.height
    if ray_table is None:
        ray_table = O.make_ray_table(w, h)
    ray_table = np.ascontiguousarray(ray_table, np.float32)
    vol = scene.volume if (scene.volume.dtype == np.int8 and scene.volume.flags.c_contiguous) else np.ascontiguousarray(scene.volume, np.int8)
    dims = (C.c_int * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    lights = np.ascontiguousarray(scene.lights, np.float32).copy()
    atlas = np.ascontiguousarray(scene.atlas, np.uint8).copy()
    rgba = np.empty((h, w, 4), np.uint8)
    rgba[...] = (255, 255, 255, 100)
    written = np.zeros((h, w), np.uint8)
    if octree is not None:
        desc = np.ascontiguousarray(octree[0], np.uint64).copy()
        root = int(octree[1])
    else:
        desc, root = np.zeros(1, np.uint64), 0
    lookup = np.zeros(max(desc.size, 1), np.uint32)
    attach = np.zeros(max(desc.size, 1), np.uint64)
    settings = np.zeros(64, np.uint64)
    settings[0], settings[1], settings[2] = scene.n, octenabled, root        # OCTDIM, OCTENABLED, OCTREE_ROOT_INDEX
    cam_dir = np.ascontiguousarray(scene.cam_dir, np.float32)
    cam_pos = np.ascontiguousarray(scene.cam_pos, np.float32)
    L = lib(lifted)
    if lifted:
        L.ref_set_max_distance(C.c_int(int(scene.max_distance)))
    else:
        assert L.ref_kernel_max_distance() == 20
    fp = C.POINTER(C.c_float)
    rc = L.ref_raycast(C.c_int(w), C.c_int(h), ray_table.ctypes.data_as(fp), vol.ctypes.data_as(C.c_void_p), dims,
                       cam_dir.ctypes.data_as(fp), cam_pos.ctypes.data_as(fp), lights.ctypes.data_as(fp), C.c_int(lights.shape[0]),
                       rgba.ctypes.data_as(C.c_void_p), written.ctypes.data_as(C.c_void_p), atlas.ctypes.data_as(C.c_void_p),
                       C.c_int(atlas.shape[1]), C.c_int(atlas.shape[0]), C.c_int(scene.tile), C.c_int(scene.tile),
                       desc.ctypes.data_as(C.c_void_p), lookup.ctypes.data_as(C.c_void_p), attach.ctypes.data_as(C.c_void_p),
                       settings.ctypes.data_as(C.c_void_p), C.c_int(rows[0] if rows else 0), C.c_int(rows[1] if rows else h), C.c_int(row_stride))
    if rc != 0:
        raise RuntimeError(f"ref_raycast failed: {rc}")
    return rgba, written.astype(bool)


# ---- the reference's host-side octree code (src/map/Octree.cpp, include/util.hpp), oracle/_ref/libref_octree.so -----
def octree_available() -> bool:
    return (REF_DIR / "libref_octree.so").exists()


def _olib() -> C.CDLL:
    if "octree" not in _libs:
        L = C.CDLL(str(REF_DIR / "libref_octree.so"))
        L.ref_octree_new.restype = C.c_void_p
        L.ref_octree_buffer_size.restype = C.c_int
        L.ref_octree_get_voxel.restype = C.c_int
        L.ref_octree_validate.restype = C.c_int
        _libs["octree"] = L
    return _libs["octree"]


class RefOctree:
    """The reference's `Octree` object: Generate / GetVoxel / Validate run the reference's own compiled code."""

    def __init__(self, volume: np.ndarray):
        import os
        import tempfile

        L = _olib()
        self.n = int(volume.shape[0])
        self.volume = np.ascontiguousarray(volume, np.int8).copy()
        self.handle = C.c_void_p(L.ref_octree_new())
        size = L.ref_octree_buffer_size()
        self.descriptors = np.zeros(size, np.uint64)
        root, pos = C.c_uint64(0), C.c_uint64(0)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as scratch:      # Generate dumps raw_output.txt / raw_data.txt (Octree.cpp:33-41)
            os.chdir(scratch)
            try:
                L.ref_octree_generate(self.handle, self.volume.ctypes.data_as(C.c_void_p), C.c_int(self.n),
                                      self.descriptors.ctypes.data_as(C.c_void_p), C.byref(root), C.byref(pos))
            finally:
                os.chdir(cwd)
        self.root_index, self.buffer_position = int(root.value), int(pos.value)

    def get_voxel(self, x: int, y: int, z: int):
        pos = (C.c_int * 3)()
        depth = C.c_int(0)
        found = _olib().ref_octree_get_voxel(self.handle, C.c_int(x), C.c_int(y), C.c_int(z), pos, C.byref(depth))
        return int(found), (pos[0], pos[1], pos[2]), int(depth.value)

    def validate(self) -> bool:
        return bool(_olib().ref_octree_validate(self.handle, self.volume.ctypes.data_as(C.c_void_p), C.c_int(self.n)))

    def close(self) -> None:
        if self.handle:
            _olib().ref_octree_free(self.handle)
            self.handle = None


def normalize(v) -> np.ndarray:
    a = np.ascontiguousarray(v, np.float32)
    out = np.zeros(3, np.float32)
    _olib().ref_normalize(a.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def viewport_available() -> bool:
    return (REF_DIR / "libref_viewport.so").exists()


def create_viewport_table(width: int, height: int) -> np.ndarray:
    """The ray table as written by the loop of the reference's CLCaster::create_viewport (src/CLCaster.cpp:244-275, cut out
    of the source and compiled by `make -C oracle ref`): float32[height, width, 4]."""
    name = "libref_viewport.so"
    if name not in _libs:
        _libs[name] = C.CDLL(str(REF_DIR / name))
    out = np.zeros((height, width, 4), np.float32)
    _libs[name].ref_create_viewport_table(C.c_int(width), C.c_int(height), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out
