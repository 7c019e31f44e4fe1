"""Deterministic synthetic scenes: procedural voxel maps, cameras, lights, atlas.

The reference builds its map on the CPU (`Map::Map`, ref src/map/Map.cpp:5-19; HEAD fills a 16^3
volume with type 5, ref src/map/ArrayMap.cpp:17-23) and loads `assets/textures/minecraft_tiles.png`
(ref src/Application.cpp:77-79).  Neither is available on the GPU box, so benchmarks and parity
tests use the seeded generators below.  Everything here is input synthesis shared by the CUDA path
and by the tests' oracle; it does no ray casting.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED_MAP = 0x5EED0001
SEED_CAMERA = 0x5EED1000


# ----------------------------------------------------------------------------- hashing / noise
def lowbias32(x: np.ndarray) -> np.ndarray:
    """32-bit integer hash (lowbias32), vectorised; wraps modulo 2^32."""
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def splitmix64(state: int) -> tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def _lattice(ix: np.ndarray, iy: np.ndarray, octave: int, seed: int) -> np.ndarray:
    h = lowbias32(ix.astype(np.uint64) * 0x9E3779B1 + lowbias32(iy.astype(np.uint64) * 0x85EBCA77 + octave * 0xC2B2AE3D + seed))
    return h.astype(np.float64) / 4294967296.0


def fbm4(n: int, seed: int = SEED_MAP) -> np.ndarray:
    """4-octave value noise on the n x n grid, in [0, 1).  Returns array [y, x]."""
    ys, xs = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    total = np.zeros((n, n), dtype=np.float64)
    amp, norm = 1.0, 0.0
    for octave in range(4):
        freq = 4 * (1 << octave)
        u = xs * freq / n
        v = ys * freq / n
        iu = np.floor(u).astype(np.int64)
        iv = np.floor(v).astype(np.int64)
        fu = u - iu
        fv = v - iv
        su = fu * fu * (3.0 - 2.0 * fu)
        sv = fv * fv * (3.0 - 2.0 * fv)
        a = _lattice(iu, iv, octave, seed)
        b = _lattice(iu + 1, iv, octave, seed)
        c = _lattice(iu, iv + 1, octave, seed)
        d = _lattice(iu + 1, iv + 1, octave, seed)
        total += amp * ((a + (b - a) * su) + ((c + (d - c) * su) - (a + (b - a) * su)) * sv)
        norm += amp
        amp *= 0.5
    return total / norm


def heightfield(n: int, seed: int = SEED_MAP) -> np.ndarray:
    """h[y, x] = clamp(floor(n * (0.25 + 0.20 * fbm4)), 1, n - 2), int32."""
    h = np.floor(n * (0.25 + 0.20 * fbm4(n, seed))).astype(np.int32)
    return np.clip(h, 1, n - 2)


def terrain_columns(n: int, variant: str = "shell", seed: int = SEED_MAP) -> tuple[np.ndarray, np.ndarray]:
    """Column description (lo, hi) int32 [y, x] of terrain_map(n, variant): voxel (x, y, z) is 5 for lo <= z <= hi."""
    h = heightfield(n, seed)
    if variant == "solid":
        lo = np.zeros_like(h)
    elif variant == "shell":
        hp = np.pad(h, 1, mode="edge")
        nb = np.minimum(np.minimum(hp[:-2, 1:-1], hp[2:, 1:-1]), np.minimum(hp[1:-1, :-2], hp[1:-1, 2:]))
        lo = np.minimum(h - 1, nb)
    else:
        raise ValueError(variant)
    return np.maximum(lo, 0).astype(np.int32), h.astype(np.int32)


def terrain_map(n: int, variant: str = "shell", seed: int = SEED_MAP, reflect_fraction: float = 0.0) -> np.ndarray:
    """Dense voxel volume as int8 array [z, y, x] (flat index x + n*(y + n*z), ref ArrayMap.cpp:39-46).

    variant "solid": voxel = 5 for z <= h(x, y).
    variant "shell": voxel = 5 for min(h - 1, lowest 4-neighbour height) <= z <= h  (watertight surface).
    reflect_fraction > 0 turns that share of the solid voxels into type 6 (reflection tests only).
    """
    h = heightfield(n, seed)
    vol = np.zeros((n, n, n), dtype=np.int8)
    if variant == "solid":
        lo = np.zeros_like(h)
    elif variant == "shell":
        hp = np.pad(h, 1, mode="edge")
        nb = np.minimum(np.minimum(hp[:-2, 1:-1], hp[2:, 1:-1]), np.minimum(hp[1:-1, :-2], hp[1:-1, 2:]))
        lo = np.minimum(h - 1, nb)
    else:
        raise ValueError(variant)
    slab = max(1, min(n, (1 << 24) // (n * n)))
    for z0 in range(0, n, slab):
        z = np.arange(z0, min(n, z0 + slab), dtype=np.int32)[:, None, None]
        vol[z0 : z0 + slab] = np.where((z <= h[None]) & (z >= lo[None]), 5, 0).astype(np.int8)
    if reflect_fraction > 0.0:
        zz, yy, xx = np.nonzero(vol)
        key = lowbias32(xx.astype(np.uint64) + n * (yy.astype(np.uint64) + n * zz.astype(np.uint64)) + (seed + 1))
        pick = key < int(reflect_fraction * 4294967296.0)
        vol[zz[pick], yy[pick], xx[pick]] = 6
    return vol


def synthetic_atlas(size: int = 256, tile: int = 16) -> np.ndarray:
    """RGBA8 atlas [y, x, 4] standing in for assets/textures/minecraft_tiles.png (256x256, 16 px tiles)."""
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    t = (x // tile) + (size // tile) * (y // tile)
    out = np.empty((size, size, 4), dtype=np.uint8)
    out[..., 0] = (x * 7 + y * 13 + t * 31) & 255
    out[..., 1] = (x * 3 + y * 5 + t * 17) & 255
    out[..., 2] = (x * 11 + y * 2 + t * 7) & 255
    out[..., 3] = 255
    return out


# ----------------------------------------------------------------------------- scene container
@dataclass
class Scene:
    """Everything `CLCaster` is given before `validate()` (ref src/Application.cpp:27-88)."""

    n: int
    volume: np.ndarray | None              # int8 [z, y, x]; None when only `columns` is given (4096^3)
    width: int
    height: int
    cam_pos: np.ndarray                    # float32[3]
    cam_dir: np.ndarray                    # float32[2]  (inclination, azimuth)
    lights: np.ndarray                     # float32[count, 10]
    atlas: np.ndarray = field(default_factory=synthetic_atlas)
    tile: int = 16
    max_distance: int = 20                 # kernel:326
    name: str = ""
    columns: tuple | None = None           # (lo, hi) int32 [y, x] when the volume is given as solid z-ranges per column


def make_lights(n: int, count: int = 1) -> np.ndarray:
    """Light slots {r,g,b,i,x,y,z,dx,dy,dz} (ref include/LightController.h:63-73); SURVEY 8d values."""
    lights = np.zeros((max(count, 1), 10), dtype=np.float32)
    lights[0] = [0.6, 0.6, 0.6, 1.0, 0.6 * n, 0.4 * n, 0.8 * n, -1.0, -1.0, -1.5]
    if count > 1:
        lights[1] = [0.6, 0.6, 0.6, 1.0, 0.2 * n, 0.7 * n, 0.9 * n, -1.0, -1.0, -1.5]
    return lights


def make_camera(n: int, h: np.ndarray, index: int = 0, need_zero_bias: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """Camera standing on the terrain: position (x, y, h(x,y)+1+frac), fractional parts in (0.05, 0.95),
    inclination in [1.2, 2.0], azimuth in [0, 2 pi).  With need_zero_bias the camera voxel shares its
    2^3 octree cell with the ground voxel below it, so the get_oct_vox start bias (kernel:353) is 0."""
    state = SEED_CAMERA + index
    for _ in range(4096):
        vals = []
        for _ in range(7):
            state, r = splitmix64(state)
            vals.append(r / 18446744073709551616.0)
        x = int(vals[0] * (n - 2)) + 1
        y = int(vals[1] * (n - 2)) + 1
        ground = int(h[y, x])
        if need_zero_bias and (ground & 1):       # cell {ground, ground+1} needs an even ground height
            continue
        if ground + 1 >= n:
            continue
        pos = np.array([x + 0.05 + 0.9 * vals[2], y + 0.05 + 0.9 * vals[3], ground + 1 + 0.05 + 0.9 * vals[4]], dtype=np.float32)
        direction = np.array([1.2 + 0.8 * vals[5], 2.0 * math.pi * vals[6]], dtype=np.float32)
        return pos, direction
    raise RuntimeError("no camera position found")


def features_map(n: int = 32) -> np.ndarray:
    """Small hand-made volume exercising every branch of the kernel: ground (5), towers that cast
    shadows, an overhang, a mirror wall and floor patch (6), transparent filler values (1, 7) and a
    large empty region above (collapsed octree cells => non-zero get_oct_vox bias for high cameras)."""
    vol = np.zeros((n, n, n), dtype=np.int8)
    g = n // 4
    vol[:g, :, :] = 5                                   # ground slab z < g
    vol[g - 1, n // 2 :, : n // 3] = 6                  # mirror floor patch
    vol[g : g + n // 3, n // 4, n // 4] = 5             # thin tower
    vol[g : g + n // 4, n // 2 : n // 2 + 2, n // 2 : n // 2 + 2] = 5   # thick tower
    vol[g + n // 4, n // 2 - 2 : n // 2 + 4, n // 2 - 2 : n // 2 + 4] = 5   # its overhanging cap
    vol[g : g + n // 2, n - 3, 2 : n - 2] = 6           # mirror wall near y = n-3
    vol[g : g + 3, 3, n // 2 :] = 5                     # low wall
    vol[g + 1, 6:9, 6:9] = 1                            # transparent values: not 5/6 => never hit
    vol[g + 2, 10, 10] = 7
    return vol


def make_scene(config: str, variant: str | None = None, lights: int = 1, camera_index: int = 0) -> Scene:
    """Named BASELINE.json configurations (sizes only)."""
    table = {
        # name: (n, width, height, default variant)
        "head": (16, 50, 50, "full"),
        "tiny": (16, 64, 48, "shell"),
        "small": (32, 160, 96, "shell"),
        "c1": (64, 1280, 720, "shell"),
        "c2": (256, 1920, 1080, "shell"),
        "c3": (1024, 3840, 2160, "shell"),
        "c5": (1024, 1920, 1080, "shell"),
    }
    if config.startswith("features"):
        # features / features-high / features-mirror: fixed cameras over features_map(32)
        n = 32
        cams = {
            "features": ([5.3, 9.7, 14.2], [2.1, 0.9]),
            "features-low": ([20.4, 5.6, 9.35], [1.65, 1.9]),
            "features-high": ([13.6, 11.2, 27.4], [2.6, 0.4]),       # camera in a collapsed empty cell
            "features-mirror": ([9.45, 20.3, 10.6], [1.7, 1.45]),     # looks at the mirror wall
        }
        pos, direction = cams[config]
        lights_arr = np.zeros((8, 10), dtype=np.float32)
        lights_arr[0] = [0.7, 0.6, 0.5, 1.0, 25.3, 6.2, 21.7, -1.0, -1.0, -1.5]
        return Scene(n, features_map(n), 200, 120, np.array(pos, np.float32), np.array(direction, np.float32),
                     lights_arr, max_distance=3 * n, name=config)
    n, w, hgt, default_variant = table[config]
    variant = variant or default_variant
    if config == "head":
        # HEAD defaults: all-5 volume, fixed camera and light, max_distance 20
        vol = np.full((n, n, n), 5, dtype=np.int8)
        lights_arr = np.zeros((8, 10), dtype=np.float32)
        lights_arr[0] = [0.01, 0.01, 0.01, 0.2, 10.0, 10.0, 10.0, -1.0, -1.0, -1.5]
        return Scene(n, vol, w, hgt, np.array([2.34, 2.5, 7.17], np.float32), np.array([2.424, 3.141], np.float32),
                     lights_arr, max_distance=20, name="head")
    vol = terrain_map(n, variant)
    h = heightfield(n)
    pos, direction = make_camera(n, h, camera_index)
    return Scene(n, vol, w, hgt, pos, direction, make_lights(n, lights), max_distance=3 * n, name=f"{config}-{variant}")
