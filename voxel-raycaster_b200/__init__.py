"""voxel-raycaster_b200 -- B200-native per-pixel voxel ray caster behind the reference's CLCaster API.

The directory name contains a hyphen, so import it with
    pkg = importlib.import_module("voxel-raycaster_b200")
(tests/conftest.py and bench.py do exactly that).  Contents:
    csrc/       hand-written sm_100a CUDA kernels + the C ABI (include/vr_caster.h) -> libvrcaster.so
    caster.py   ctypes binding and `CUDACaster`, the host-side mirror of CLCaster
    scene.py    deterministic synthetic scenes (maps, cameras, lights, atlas)
"""
from . import scene, tiles  # noqa: F401
from .caster import AUX_DTYPE, SYMBOLS, CUDACaster, VrStats, load_library, octree_generate, octree_get_voxel  # noqa: F401
