"""CPU experiment (no GPU): how many turns of its cell loop does the closed-form walk (csrc/vr_canon.h) make on a bench frame,
by kind and size of the empty cell it is handed -- over the undirected top grid, over the directed grids, and over
directed grids whose cubes are grown anisotropically (an experiment that was NOT adopted: DESIGN.md section 4).

The device core is compiled for the host (tests/host_emu/emu.cpp) with the analysis hook VR_CANON_STAT defined; the
anisotropic variant needs three small edits of vr_canon.h, which are applied to a TEMPORARY COPY of the sources (the
product headers stay as they are).  Prints, per grid, turns per pixel for primary and shadow rays by cell kind, the mean over
8x4-pixel warps of the maximum lookup count (the number of loop turns a warp executes), and checks that the frames are equal.

    python profiles/canon_stats.py [config=c3] [tile_stride=16]
"""
import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import emu_lib  # noqa: E402
import oracle_lib as O  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "c3"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 16

PRE = r'''
#include <stdint.h>
#include <stdlib.h>
#include <vector>
extern "C" { extern double g_stat[2][3][64][2]; }   /* [shadow][kind: 0 brick, 1 cube (ext > 0), 2 aligned cell][size bucket][turns, steps] */
static thread_local int g_before;
static inline void canon_stat(int kind, int m, int ext, int steps, bool shadow) {
    if (kind == 0) { g_before = steps; return; }
    int k, b;
    if (kind == 1) { k = 0; b = steps > 63 ? 63 : steps; }
    else { steps += g_before; if (ext > 0) { k = 1; b = ext > 63 ? 63 : ext; } else { k = 2; int l = 0; while ((1 << l) - 1 < m) l++; b = l; } }
    #pragma omp atomic
    g_stat[shadow][k][b][0] += 1;
    #pragma omp atomic
    g_stat[shadow][k][b][1] += steps;
}
#define VR_CANON_STAT(brick, m, ext, steps, shadow) canon_stat(brick, m, ext, steps, shadow)
'''

ANISO = r'''
double g_stat[2][3][64][2];
static std::vector<int32_t> g_sat[8];
static int g_G, g_mode, g_cap, g_maxe;
extern "C" void aniso_setup(const uint32_t *grid, int G) {          /* summed-area tables of the non-empty blocks, mirrored per octant */
    g_G = G;
    g_mode = getenv("ANISO") ? atoi(getenv("ANISO")) : 0;
    g_cap = getenv("ANISO_CAP") ? atoi(getenv("ANISO_CAP")) : 64;
    g_maxe = getenv("ANISO_MAXE") ? atoi(getenv("ANISO_MAXE")) : 1000;
    if (!g_mode) return;
    const size_t n3 = (size_t)G * G * G;
#pragma omp parallel for
    for (int o = 0; o < 8; o++) {
        auto &S = g_sat[o];
        S.assign(n3, 0);
        auto at = [&](int x, int y, int z) -> int32_t { return (x < 0 || y < 0 || z < 0) ? 0 : S[(size_t)x + (size_t)G * (y + (size_t)G * z)]; };
        for (int kz = 0; kz < G; kz++) for (int ky = 0; ky < G; ky++) for (int kx = 0; kx < G; kx++) {
            const int bx = (o & 1) ? G - 1 - kx : kx, by = (o & 2) ? G - 1 - ky : ky, bz = (o & 4) ? G - 1 - kz : kz;
            const int occ = (grid[(size_t)bx + (size_t)G * (by + (size_t)G * bz)] >> 31) & 1;
            S[(size_t)kx + (size_t)G * (ky + (size_t)G * kz)] = occ + at(kx-1,ky,kz) + at(kx,ky-1,kz) + at(kx,ky,kz-1) - at(kx-1,ky-1,kz) - at(kx-1,ky,kz-1) - at(kx,ky-1,kz-1) + at(kx-1,ky-1,kz-1);
        }
    }
}
static int box_count(int o, int x0, int y0, int z0, int x1, int y1, int z1) {
    const int G = g_G; auto &S = g_sat[o];
    auto at = [&](int x, int y, int z) -> int32_t { return (x < 0 || y < 0 || z < 0) ? 0 : S[(size_t)x + (size_t)G * (y + (size_t)G * z)]; };
    return at(x1,y1,z1) - at(x0-1,y1,z1) - at(x1,y0-1,z1) - at(x1,y1,z0-1) + at(x0-1,y0-1,z1) + at(x0-1,y1,z0-1) + at(x1,y0-1,z0-1) - at(x0-1,y0-1,z0-1);
}
/* greedy growth of the cube of edge E at mirrored block k: one more layer per axis, round robin, while the slab is empty */
extern "C" void aniso_box(int o, int kx, int ky, int kz, int E, int *ax, int *ay, int *az) {
    *ax = *ay = *az = 0;
    if (!g_mode || E > g_maxe) return;
    int ex = E, ey = E, ez = E;
    const int G = g_G;
    for (bool any = true; any;) {
        any = false;
        if (ex < g_cap && kx + ex < G && box_count(o, kx + ex, ky, kz, kx + ex, ky + ey - 1, kz + ez - 1) == 0) { ex++; any = true; }
        if (ey < g_cap && ky + ey < G && box_count(o, kx, ky + ey, kz, kx + ex - 1, ky + ey, kz + ez - 1) == 0) { ey++; any = true; }
        if (ez < g_cap && kz + ez < G && box_count(o, kx, ky, kz + ez, kx + ex - 1, ky + ey - 1, kz + ez) == 0) { ez++; any = true; }
    }
    *ax = ex - E; *ay = ey - E; *az = ez - E;
}
'''


def patched(text: str, pairs) -> str:
    for old, new in pairs:
        assert old in text, old
        text = text.replace(old, new)
    return text


tmp = Path(tempfile.mkdtemp(prefix="vr_canon_stats."))
(tmp / "csrc").mkdir()
(tmp / "emu").mkdir()
for f in (ROOT / "voxel-raycaster_b200" / "csrc").glob("*.h"):
    shutil.copy(f, tmp / "csrc")
shutil.copy(ROOT / "voxel-raycaster_b200" / "csrc" / "vr_octree.cpp", tmp / "csrc")
canon = patched((tmp / "csrc" / "vr_canon.h").read_text(), [
    ("struct vr_ccell {\n    int m, ext;\n    bool brick;\n};",
     "struct vr_ccell {\n    int m, ext;\n    bool brick;\n    int ax = 0, ay = 0, az = 0;\n};\n"
     "extern \"C\" void aniso_box(int o, int kx, int ky, int kz, int E, int *ax, int *ay, int *az);"),
    ("            c.ext = (int)(e >> 8);\n            c.brick = false;\n            return 0;",
     "            c.ext = (int)(e >> 8);\n            c.brick = false;\n            c.ax = c.ay = c.az = 0;\n"
     "            if ((int)(e & 31u) == g && P.grid_directed) {\n"
     "                aniso_box((int)(q.gflip >> (3 * P.grid_bits)), q.px >> g, q.py >> g, q.pz >> g, (c.ext >> g) + 1, &c.ax, &c.ay, &c.az);\n"
     "                c.ax <<= g; c.ay <<= g; c.az <<= g;\n            }\n            return 0;"),
    ("            c.ext = 0;\n            return 0;", "            c.ext = 0; c.ax = c.ay = c.az = 0;\n            return 0;"),
    ("const int ox = (q.px | c.m) + c.ext, oy = (q.py | c.m) + c.ext, oz = (q.pz | c.m) + c.ext;",
     "const int ox = (q.px | c.m) + c.ext + c.ax, oy = (q.py | c.m) + c.ext + c.ay, oz = (q.pz | c.m) + c.ext + c.az;"),
])
(tmp / "csrc" / "vr_canon.h").write_text(canon)
emu = patched((ROOT / "tests" / "host_emu" / "emu.cpp").read_text(), [
    ("../../voxel-raycaster_b200/csrc/", "../csrc/"),
    ("        P.grid_directed = use_svo == 4;", "        P.grid_directed = use_svo == 4;\n        if (use_svo == 4) aniso_setup(grid.data(), 1 << gb);"),
    ('extern "C" int emu_raycast(', 'extern "C" void aniso_setup(const uint32_t *grid, int G);\nextern "C" int emu_raycast('),
])
(tmp / "emu" / "emu.cpp").write_text(emu)
(tmp / "stats.cpp").write_text(PRE + ANISO + '#include "emu/emu.cpp"\nextern "C" double *emu_stats() { return &g_stat[0][0][0][0]; }\n')
subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", f"-I{tmp}",
                "-o", str(tmp / "libstats.so"), str(tmp / "stats.cpp"), str(tmp / "csrc" / "vr_octree.cpp")], check=True)

scene = bench.bench_scene(config)
rows = np.concatenate([np.arange(r, r + 4) for r in range(0, scene.height, 4 * stride)])      # whole 8x4 warp tiles
table = O.make_ray_table(scene.width, scene.height)[rows].copy()
sub = scene.__class__(scene.n, scene.volume, scene.width, len(rows), scene.cam_pos, scene.cam_dir, scene.lights, atlas=scene.atlas,
                      max_distance=scene.max_distance)
names = ["brick", "cube", "aligned cell"]
frames = []
for label, use_svo, env in (("undirected top grid", 3, {}), ("directed top grids", 4, {}),
                            ("directed grids + anisotropic growth (experiment)", 4, {"ANISO": "1"}),
                            ("directed grids + anisotropic growth of cubes up to 8 blocks, at most 24 (experiment)", 4,
                             {"ANISO": "1", "ANISO_MAXE": "8", "ANISO_CAP": "24"})):
    for k in ("ANISO", "ANISO_MAXE", "ANISO_CAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    # a fresh copy of the library per run: its counters and getenv-initialised settings are process-wide statics
    so = tmp / f"libstats_{len(frames)}.so"
    shutil.copy(tmp / "libstats.so", so)
    lib = C.CDLL(str(so))
    lib.emu_raycast.restype = C.c_int
    emu_lib._lib = lib
    rgba, aux = emu_lib.raycast(sub, table, use_svo=use_svo)
    lib.emu_stats.restype = C.POINTER(C.c_double)
    st = np.ctypeslib.as_array(lib.emu_stats(), shape=(2, 3, 64, 2)).copy()
    npx = rgba.shape[0] * rgba.shape[1]
    lk = aux["lookups"].astype(np.int64)
    H, W = lk.shape
    warps = lk.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(H // 4, W // 8, 32)
    print(f"== {label}: lookups per pixel {lk.mean():.2f}, loop turns per warp (mean of the per-warp maximum) {warps.max(-1).mean():.2f}")
    for sh in (0, 1):
        parts = ", ".join(f"{names[k]} {st[sh, k, :, 0].sum() / npx:.2f} ({st[sh, k, :, 1].sum() / max(st[sh, k, :, 0].sum(), 1):.1f} steps each)" for k in range(3))
        print(f"   {'shadow ' if sh else 'primary'} ray: {st[sh, :, :, 0].sum() / npx:.2f} turns per pixel: {parts}")
    frames.append(rgba)
for f in frames[1:]:
    assert np.array_equal(f, frames[0]), "the frame must not depend on the grid (tie rays aside)"
print("frames identical on the sampled rows")
shutil.rmtree(tmp, ignore_errors=True)
