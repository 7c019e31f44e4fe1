"""ctypes binding of libvrcaster.so and a host-side mirror of the reference's `CLCaster` class.

`CUDACaster` keeps the reference's method names, call order and bool-return convention
(ref include/CLCaster.h:110-179): init, create_viewport, assign_lights, assign_map, assign_octree,
assign_camera, create_texture_atlas, validate, compute, draw, add_to_settings_buffer,
overwrite_setting, ...  SFML types are replaced by numpy arrays; `draw` returns the frame instead of
blitting a sprite (headless mode).  Every call goes through the C ABI in include/vr_caster.h; there is
no Python or CPU implementation behind it -- importing works without a GPU, `init()` does not.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

# VR_CASTER_LIB: another build of the same library (kernel A/B runs on the GPU box); default = the in-tree build
_LIB_PATH = Path(os.environ.get("VR_CASTER_LIB") or Path(__file__).resolve().parent / "libvrcaster.so")
_lib = None

# every symbol include/vr_caster.h declares: name -> (restype, argtypes)
_vp, _i, _u64p, _f32p, _u8p, _i8p, _i64p, _i32p = C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int64), C.POINTER(C.c_int32)


class VrStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("frames", C.c_uint64), ("native_nodes", C.c_uint64),
        ("native_bytes", C.c_uint64), ("solid_voxels", C.c_uint64), ("levels", C.c_int32),
        ("used_svo", C.c_int32), ("bias", C.c_int32 * 3), ("device", C.c_int32), ("last_kernel_ms", C.c_float),
        ("build_ms", C.c_float), ("build_masks_ms", C.c_float),
    ]


AUX_DTYPE = np.dtype([
    ("hit", "<i4", (3,)), ("face", "u1"), ("status", "u1"), ("flags", "u1"), ("hit_type", "u1"),
    ("steps_first", "<u4"), ("steps_total", "<u4"), ("node_fetches", "<u4"), ("lookups", "<u4"),
])
assert AUX_DTYPE.itemsize == 32

SYMBOLS = {
    "vr_init": (_i, [C.POINTER(_vp), _i, C.c_uint]),
    "vr_destroy": (None, [_vp]),
    "vr_last_error": (C.c_char_p, [_vp]),
    "vr_version": (C.c_char_p, []),
    "vr_load_config": (_i, [_vp, C.c_char_p]),
    "vr_save_config": (_i, [_vp, C.c_char_p]),
    "vr_create_viewport": (_i, [_vp, _i, _i, C.c_float, C.c_float]),
    "vr_release_viewport": (_i, [_vp]),
    "vr_assign_lights": (_i, [_vp, _f32p, _i]),
    "vr_assign_map": (_i, [_vp, _i8p, _i, _i, _i]),
    "vr_release_map": (_i, [_vp]),
    "vr_assign_columns": (_i, [_vp, _i32p, _i32p, _i, _i]),
    "vr_assign_octree": (_i, [_vp, _u64p, C.POINTER(C.c_uint32), _u64p, C.c_uint64, C.c_uint64]),
    "vr_release_octree": (_i, [_vp]),
    "vr_assign_camera": (_i, [_vp, _f32p, _f32p]),
    "vr_release_camera": (_i, [_vp]),
    "vr_create_texture_atlas": (_i, [_vp, _u8p, _i, _i, _i, _i]),
    "vr_create_settings_buffer": (_i, [_vp]),
    "vr_release_settings_buffer": (_i, [_vp]),
    "vr_add_to_settings_buffer": (_i, [_vp, C.c_char_p, C.c_char_p, C.c_int64]),
    "vr_overwrite_setting": (_i, [_vp, C.c_char_p, _i64p]),
    "vr_remove_from_settings_buffer": (_i, [_vp, C.c_char_p]),
    "vr_settings_data": (_i64p, [_vp]),
    "vr_set_define": (_i, [_vp, C.c_char_p, C.c_char_p]),
    "vr_remove_define": (_i, [_vp, C.c_char_p]),
    "vr_validate": (_i, [_vp]),
    "vr_debug_quick_recompile": (_i, [_vp]),
    "vr_compute": (_i, [_vp]),
    "vr_compute_async": (_i, [_vp]),
    "vr_sync": (_i, [_vp]),
    "vr_compute_into": (_i, [_vp, _vp]),
    "vr_compute_views": (_i, [_vp, _f32p, _i, _vp]),
    "vr_read_framebuffer": (_i, [_vp, _u8p, C.c_size_t]),
    "vr_frame_begin": (_i, [_vp]),
    "vr_frame_end": (_i, [_vp, C.POINTER(_u8p)]),
    "vr_set_bands": (_i, [_vp, _i, _i, _i]),
    "vr_local_rows": (_i, [_vp]),
    "vr_set_tiles": (_i, [_vp, _i, _i]),
    "vr_set_option": (_i, [_vp, C.c_char_p, C.c_int64]),
    "vr_set_stream": (_i, [_vp, _vp]),
    "vr_enable_aux": (_i, [_vp, _i]),
    "vr_read_aux": (_i, [_vp, _vp, C.c_size_t]),
    "vr_device_image": (_vp, [_vp]),
    "vr_read_ray_table": (_i, [_vp, _f32p, C.c_size_t]),
    "vr_native_tree_info": (_i, [_vp, _u64p, _u64p, _i32p, _i32p]),
    "vr_native_tree_copy": (_i, [_vp, _vp, _vp]),
    "vr_gl_register_texture": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "vr_gl_draw": (C.c_int, [_vp]),
    "vr_gl_unregister": (C.c_int, [_vp]),
    "vr_top_grid_read": (C.c_uint64, [_vp, _vp, C.c_uint64, _i32p, _i32p]),
    "vr_mgpu_init": (_i, [_vp, C.c_char_p, _i, _i, C.c_uint]),
    "vr_mgpu_broadcast_octree": (_i, [_vp]),
    "vr_mgpu_frame": (_i, [_vp, _u64p]),
    "vr_mgpu_frame_wait": (_i, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "vr_mgpu_frame_release": (_i, [_vp, C.c_uint64]),
    "vr_mgpu_flush": (_i, [_vp]),
    "vr_mgpu_barrier": (_i, [_vp]),
    "vr_mgpu_shutdown": (_i, [_vp]),
    "vr_assign_native_tree": (_i, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.c_int32, C.c_int32]),
    "vr_device_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "vr_device_free": (_i, [_vp, _vp]),
    "vr_ipc_get_handle": (_i, [_vp, _vp, _vp]),
    "vr_ipc_open_handle": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "vr_ipc_close_handle": (_i, [_vp, _vp]),
    "vr_push_bands": (_i, [_vp, _vp, _vp, _vp]),
    "vr_host_register": (_i, [_vp, _vp, C.c_size_t]),
    "vr_host_unregister": (_i, [_vp, _vp]),
    "vr_octree_save": (_i, [_vp, C.c_char_p]),
    "vr_octree_load": (_i, [_vp, C.c_char_p]),
    "vr_get_stats": (_i, [_vp, C.POINTER(VrStats)]),
    "vr_octree_generate": (_i, [_i8p, _i, _u64p, _u64p, _u64p]),
    "vr_octree_get_voxel": (_i, [_u64p, C.c_uint64, C.c_uint64, _i, _i32p, _i32p, _i32p]),
}


def load_library() -> C.CDLL:
    """Loads libvrcaster.so (built in-tree by `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the caster has no CPU implementation)")
        lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def octree_generate(volume: np.ndarray) -> tuple[np.ndarray, int]:
    """`Octree::Generate` equivalent (ref src/map/Octree.cpp:13): reference-format descriptor buffer."""
    lib = load_library()
    vol = np.ascontiguousarray(volume, dtype=np.int8)
    n = vol.shape[0]
    entries, root = C.c_uint64(0), C.c_uint64(0)
    if not lib.vr_octree_generate(_ptr(vol, C.c_int8), n, None, C.byref(entries), C.byref(root)):
        raise RuntimeError("vr_octree_generate failed")
    buf = np.zeros(entries.value, dtype=np.uint64)
    if not lib.vr_octree_generate(_ptr(vol, C.c_int8), n, _ptr(buf, C.c_uint64), C.byref(entries), C.byref(root)):
        raise RuntimeError("vr_octree_generate failed")
    return buf, int(root.value)


def octree_get_voxel(desc: np.ndarray, root: int, dim: int, pos) -> tuple[int, tuple[int, int, int], int]:
    lib = load_library()
    p = (C.c_int32 * 3)(*[int(v) for v in pos])
    sub = (C.c_int32 * 3)()
    res = C.c_int32(0)
    found = lib.vr_octree_get_voxel(_ptr(desc, C.c_uint64), desc.size, root, dim, p, sub, C.byref(res))
    return int(found), (sub[0], sub[1], sub[2]), int(res.value)


class CUDACaster:
    """Host-side mirror of `CLCaster` (ref include/CLCaster.h:93) over the C ABI."""

    def __init__(self) -> None:
        self._lib = load_library()
        self._ctx = _vp()
        self._keep: dict[str, np.ndarray] = {}      # arrays whose memory the library aliases
        self.width = self.height = 0

    # -- lifecycle ------------------------------------------------------------------------------
    def init(self, device: int = 0) -> bool:
        return bool(self._lib.vr_init(C.byref(self._ctx), device, 1))

    def close(self) -> None:
        if self._ctx:
            self._lib.vr_destroy(self._ctx)
            self._ctx = _vp()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return self._lib.vr_last_error(self._ctx).decode()

    # -- scene ----------------------------------------------------------------------------------
    def create_viewport(self, width: int, height: int, v_fov: float = 0.0, h_fov: float = 0.0) -> bool:
        ok = bool(self._lib.vr_create_viewport(self._ctx, width, height, v_fov, h_fov))
        if ok:
            self.width, self.height = width, height
        return ok

    def release_viewport(self) -> bool:
        return bool(self._lib.vr_release_viewport(self._ctx))

    def assign_lights(self, lights: np.ndarray) -> bool:
        """lights: float32 [count, 10]; ALIASED -- later in-place edits are seen by the next compute()."""
        assert lights.dtype == np.float32 and lights.flags.c_contiguous and lights.shape[-1] == 10
        self._keep["lights"] = lights
        return bool(self._lib.vr_assign_lights(self._ctx, _ptr(lights, C.c_float), lights.shape[0]))

    def assign_map(self, volume: np.ndarray) -> bool:
        """volume: int8 [z, y, x] (copied)."""
        vol = np.ascontiguousarray(volume, dtype=np.int8)
        nz, ny, nx = vol.shape
        return bool(self._lib.vr_assign_map(self._ctx, _ptr(vol, C.c_int8), nx, ny, nz))

    def assign_columns(self, lo: np.ndarray, hi: np.ndarray, voxel_type: int = 5) -> bool:
        """lo, hi: int32 [y, x]; column (x, y) is solid for lo <= z <= hi.  Octree traversal only."""
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        return bool(self._lib.vr_assign_columns(self._ctx, _ptr(lo, C.c_int32), _ptr(hi, C.c_int32), lo.shape[0], voxel_type))

    def release_map(self) -> bool:
        return bool(self._lib.vr_release_map(self._ctx))

    def assign_octree(self, descriptors: np.ndarray, root_index: int) -> bool:
        d = np.ascontiguousarray(descriptors, dtype=np.uint64)
        return bool(self._lib.vr_assign_octree(self._ctx, _ptr(d, C.c_uint64), None, None, d.size, root_index))

    def release_octree(self) -> bool:
        return bool(self._lib.vr_release_octree(self._ctx))

    def assign_camera(self, direction: np.ndarray, position: np.ndarray) -> bool:
        """direction float32[2], position float32[3]; both ALIASED like Camera::get_*_pointer."""
        assert direction.dtype == np.float32 and position.dtype == np.float32
        self._keep["cam_dir"], self._keep["cam_pos"] = direction, position
        return bool(self._lib.vr_assign_camera(self._ctx, _ptr(direction, C.c_float), _ptr(position, C.c_float)))

    def release_camera(self) -> bool:
        return bool(self._lib.vr_release_camera(self._ctx))

    def create_texture_atlas(self, rgba: np.ndarray, tile_dim=(16, 16)) -> bool:
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = a.shape[:2]
        return bool(self._lib.vr_create_texture_atlas(self._ctx, _ptr(a, C.c_uint8), w, h, tile_dim[0], tile_dim[1]))

    # -- settings -------------------------------------------------------------------------------
    def create_settings_buffer(self) -> bool:
        return bool(self._lib.vr_create_settings_buffer(self._ctx))

    def release_settings_buffer(self) -> bool:
        return bool(self._lib.vr_release_settings_buffer(self._ctx))

    def add_to_settings_buffer(self, setting_name: str, define_accessor_name: str, value: int) -> bool:
        return bool(self._lib.vr_add_to_settings_buffer(self._ctx, setting_name.encode(), define_accessor_name.encode(), int(value)))

    def overwrite_setting(self, setting_name: str, value: int) -> bool:
        v = C.c_int64(int(value))
        return bool(self._lib.vr_overwrite_setting(self._ctx, setting_name.encode(), C.byref(v)))

    def remove_from_settings_buffer(self, setting_name: str) -> bool:
        return bool(self._lib.vr_remove_from_settings_buffer(self._ctx, setting_name.encode()))

    def settings(self) -> np.ndarray:
        p = self._lib.vr_settings_data(self._ctx)
        return np.ctypeslib.as_array(p, shape=(64,))

    def set_define(self, name: str, value: str) -> None:
        self._lib.vr_set_define(self._ctx, name.encode(), value.encode())

    def remove_define(self, name: str) -> None:
        self._lib.vr_remove_define(self._ctx, name.encode())

    def load_config(self, path: str | None = None) -> bool:
        return bool(self._lib.vr_load_config(self._ctx, path.encode() if path else None))

    def save_config(self, path: str | None = None) -> bool:
        return bool(self._lib.vr_save_config(self._ctx, path.encode() if path else None))

    # -- per frame ------------------------------------------------------------------------------
    def validate(self) -> bool:
        return bool(self._lib.vr_validate(self._ctx))

    def debug_quick_recompile(self) -> bool:
        return bool(self._lib.vr_debug_quick_recompile(self._ctx))

    def compute(self) -> bool:
        return bool(self._lib.vr_compute(self._ctx))

    def compute_async(self) -> bool:
        return bool(self._lib.vr_compute_async(self._ctx))

    def sync(self) -> bool:
        return bool(self._lib.vr_sync(self._ctx))

    def compute_into(self, device_ptr: int) -> bool:
        return bool(self._lib.vr_compute_into(self._ctx, _vp(device_ptr)))

    def compute_views(self, cameras: np.ndarray, device_ptr: int) -> bool:
        """cameras float32 [count, 5] = {inclination, azimuth, x, y, z}; frames land consecutively at device_ptr."""
        cams = np.ascontiguousarray(cameras, dtype=np.float32)
        self._keep["views"] = cams
        return bool(self._lib.vr_compute_views(self._ctx, _ptr(cams, C.c_float), cams.shape[0], _vp(device_ptr)))

    def draw(self) -> np.ndarray:
        """Headless `draw`: the last frame as uint8 [rows, width, 4] (rows = local slab when banded)."""
        rows = self.local_rows()
        out = np.empty((rows, self.width, 4), dtype=np.uint8)
        if not self._lib.vr_read_framebuffer(self._ctx, _ptr(out, C.c_uint8), out.nbytes):
            raise RuntimeError(self.last_error())
        return out

    def frame_begin(self) -> bool:
        return bool(self._lib.vr_frame_begin(self._ctx))

    def frame_end(self) -> np.ndarray:
        p = _u8p()
        if not self._lib.vr_frame_end(self._ctx, C.byref(p)):
            raise RuntimeError(self.last_error())
        return np.ctypeslib.as_array(p, shape=(self.local_rows(), self.width, 4))

    # -- extensions -----------------------------------------------------------------------------
    def set_bands(self, band_rows: int, stride: int, first: int) -> bool:
        return bool(self._lib.vr_set_bands(self._ctx, band_rows, stride, first))

    def set_tiles(self, world: int, rank: int) -> bool:
        """2-D tile interleave: render tiles (tx + ty) % world == rank in place into a full-size frame"""
        return bool(self._lib.vr_set_tiles(self._ctx, world, rank))

    def set_option(self, name: str, value: int) -> bool:
        return bool(self._lib.vr_set_option(self._ctx, name.encode(), int(value)))

    def local_rows(self) -> int:
        return int(self._lib.vr_local_rows(self._ctx))

    def set_stream(self, cuda_stream: int | None) -> bool:
        return bool(self._lib.vr_set_stream(self._ctx, _vp(cuda_stream or 0)))

    def enable_aux(self, enable: bool = True) -> bool:
        return bool(self._lib.vr_enable_aux(self._ctx, int(enable)))

    def read_aux(self) -> np.ndarray:
        out = np.zeros((self.local_rows(), self.width), dtype=AUX_DTYPE)
        if not self._lib.vr_read_aux(self._ctx, out.ctypes.data_as(_vp), out.nbytes):
            raise RuntimeError(self.last_error())
        return out

    def read_ray_table(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        if not self._lib.vr_read_ray_table(self._ctx, _ptr(out, C.c_float), out.nbytes):
            raise RuntimeError(self.last_error())
        return out

    def device_image(self) -> int:
        return int(self._lib.vr_device_image(self._ctx) or 0)

    def native_tree_info(self) -> tuple[int, int, int, int]:
        """(node bytes, leaf-type bytes, levels, map edge) of the 64-tree."""
        nb, tb, lv, dm = C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_int32(0)
        if not self._lib.vr_native_tree_info(self._ctx, C.byref(nb), C.byref(tb), C.byref(lv), C.byref(dm)):
            raise RuntimeError(self.last_error())
        return int(nb.value), int(tb.value), int(lv.value), int(dm.value)

    # ---- multi-GPU frame scheduler (csrc/vr_mgpu.cu); thin bindings, one call each
    MGPU_HOST_FRAME = 1

    def mgpu_init(self, session: str, world: int, rank: int, flags: int = 0) -> bool:
        try:                       # the scheduler uses the NCCL copy the process has already loaded: PyTorch's, if it is around
            import torch  # noqa: F401
        except ImportError:
            pass
        return bool(self._lib.vr_mgpu_init(self._ctx, session.encode(), world, rank, flags))

    def mgpu_broadcast_octree(self) -> bool:
        return bool(self._lib.vr_mgpu_broadcast_octree(self._ctx))

    def mgpu_frame(self) -> int:
        """enqueues this rank's share of the next frame; returns the frame number (-1 on failure)"""
        k = C.c_uint64(0)
        return int(k.value) if self._lib.vr_mgpu_frame(self._ctx, C.byref(k)) else -1

    def mgpu_frame_wait(self, frame_no: int) -> int | None:
        """root: address of the assembled frame (device, or host with MGPU_HOST_FRAME); None on failure; 0 on other ranks"""
        ptr = _vp()
        if not self._lib.vr_mgpu_frame_wait(self._ctx, frame_no, C.byref(ptr)):
            return None
        return int(ptr.value or 0)

    def mgpu_frame_release(self, frame_no: int) -> bool:
        return bool(self._lib.vr_mgpu_frame_release(self._ctx, frame_no))

    def mgpu_flush(self) -> bool:
        return bool(self._lib.vr_mgpu_flush(self._ctx))

    def mgpu_barrier(self) -> bool:
        """collective, CPU only: every rank leaves within a microsecond of the others (spin on the shared segment)"""
        return bool(self._lib.vr_mgpu_barrier(self._ctx))

    def mgpu_shutdown(self) -> bool:
        return bool(self._lib.vr_mgpu_shutdown(self._ctx))

    def gl_register_texture(self, gl_texture: int, gl_target: int = 0x0DE1) -> bool:
        """CUDA-GL interop of the viewer path (CLCaster::draw with CL/GL sharing): registers an RGBA8 GL texture (default
        target GL_TEXTURE_2D); needs a current OpenGL context."""
        return bool(self._lib.vr_gl_register_texture(self._ctx, int(gl_texture), int(gl_target)))

    def gl_draw(self) -> bool:
        return bool(self._lib.vr_gl_draw(self._ctx))

    def gl_unregister(self) -> bool:
        return bool(self._lib.vr_gl_unregister(self._ctx))

    def top_grid(self) -> tuple[np.ndarray, int, int]:
        """(grid entries, block shift, log2 G) of the closed-form walk's top grid: uint32[8, G, G, G] indexed [octant, z, y, x]
        (option directed_grid = 1, the default: one table per direction octant of a ray) or uint32[G, G, G]."""
        gs, gb = C.c_int32(0), C.c_int32(0)
        n = int(self._lib.vr_top_grid_read(self._ctx, None, 0, C.byref(gs), C.byref(gb)))
        if n == 0:
            raise RuntimeError(self.last_error())
        out = np.zeros(n, dtype=np.uint32)
        if int(self._lib.vr_top_grid_read(self._ctx, out.ctypes.data_as(_vp), n, C.byref(gs), C.byref(gb))) != n:
            raise RuntimeError(self.last_error())
        g = 1 << gb.value
        return (out.reshape(8, g, g, g) if n == 8 * g ** 3 else out.reshape(g, g, g)), int(gs.value), int(gb.value)

    def native_tree_copy(self, device_nodes: int, device_types: int) -> bool:
        return bool(self._lib.vr_native_tree_copy(self._ctx, _vp(device_nodes), _vp(device_types)))

    def assign_native_tree(self, device_nodes: int, node_bytes: int, device_types: int, type_bytes: int, levels: int, dim: int) -> bool:
        return bool(self._lib.vr_assign_native_tree(self._ctx, _vp(device_nodes), node_bytes, _vp(device_types), type_bytes, levels, dim))

    def device_alloc(self, nbytes: int) -> int:
        out = _vp()
        if not self._lib.vr_device_alloc(self._ctx, nbytes, C.byref(out)):
            raise RuntimeError(self.last_error())
        return int(out.value)

    def device_free(self, device_ptr: int) -> bool:
        return bool(self._lib.vr_device_free(self._ctx, _vp(device_ptr)))

    def ipc_get_handle(self, device_ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        if not self._lib.vr_ipc_get_handle(self._ctx, _vp(device_ptr), buf):
            raise RuntimeError(self.last_error())
        return buf.raw

    def ipc_open_handle(self, handle: bytes) -> int:
        out = _vp()
        if not self._lib.vr_ipc_open_handle(self._ctx, C.c_char_p(handle), C.byref(out)):
            raise RuntimeError(self.last_error())
        return int(out.value)

    def ipc_close_handle(self, device_ptr: int) -> bool:
        return bool(self._lib.vr_ipc_close_handle(self._ctx, _vp(device_ptr)))

    def push_bands(self, slab_ptr: int, frame_ptr: int, cuda_stream: int | None = None) -> bool:
        return bool(self._lib.vr_push_bands(self._ctx, _vp(slab_ptr), _vp(frame_ptr), _vp(cuda_stream or 0)))

    def host_register(self, host_ptr: int, nbytes: int) -> bool:
        """page-lock caller-owned host memory (e.g. a shared-memory frame) so push_bands can target it"""
        return bool(self._lib.vr_host_register(self._ctx, _vp(host_ptr), nbytes))

    def host_unregister(self, host_ptr: int) -> bool:
        return bool(self._lib.vr_host_unregister(self._ctx, _vp(host_ptr)))

    def octree_save(self, path: str) -> bool:
        return bool(self._lib.vr_octree_save(self._ctx, str(path).encode()))

    def octree_load(self, path: str) -> bool:
        return bool(self._lib.vr_octree_load(self._ctx, str(path).encode()))

    def stats(self) -> VrStats:
        s = VrStats()
        self._lib.vr_get_stats(self._ctx, C.byref(s))
        return s

    # -- convenience: the reference's init order (ref src/Application.cpp:27-88) -----------------
    def load_scene(self, scene, use_octree: bool, assign_octree: bool = True, device: int = 0, shadow_lights: int = 1,
                   walk: int | None = None, gpu_build: bool | None = None, collapse_solid: bool | None = None) -> None:
        def must(ok: bool, what: str) -> None:
            if not ok:
                raise RuntimeError(f"{what} failed: {self.last_error()}")

        must(self.init(device), "init")
        if gpu_build is not None:  # None = the library default (the 64-tree is built on the device)
            must(self.set_option("gpu_build", 1 if gpu_build else 0), "set_option gpu_build")
        if collapse_solid is not None:  # None = the library default (solid subtrees of the 64-tree become single nodes)
            must(self.set_option("collapse_solid", 1 if collapse_solid else 0), "set_option collapse_solid")
        must(self.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n), "add OCTDIM")
        must(self.add_to_settings_buffer("using_octree", "OCTENABLED", 0 if use_octree else 1), "add OCTENABLED")
        must(self.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance), "add MAX_DISTANCE")
        if shadow_lights > 1:      # extension: the reference binds light_count but reads light 0 only
            must(self.add_to_settings_buffer("light_count", "LIGHT_COUNT", shadow_lights), "add LIGHT_COUNT")
        if scene.volume is None:
            must(self.assign_columns(scene.columns[0], scene.columns[1]), "assign_columns")
        else:
            if assign_octree:
                desc, root = octree_generate(scene.volume)
                must(self.assign_octree(desc, root), "assign_octree")
            must(self.assign_map(scene.volume), "assign_map")
        must(self.assign_camera(scene.cam_dir, scene.cam_pos), "assign_camera")
        must(self.create_viewport(scene.width, scene.height, 0.625 * 90.0, 90.0), "create_viewport")
        must(self.assign_lights(scene.lights), "assign_lights")
        must(self.create_texture_atlas(scene.atlas, (scene.tile, scene.tile)), "create_texture_atlas")
        must(self.validate(), "validate")
        if walk is not None:       # None = the library default (2: closed-form crossing times); 0 / 1 = the exact walks
            must(self.set_option("walk", walk), "set_option walk")
