"""Writes the inputs of profiles/quick_ab/quick_ab.cpp (a few-second GPU check without Python: the last seconds of the round's
GPU budget) into build/quick/: the C3 scene as column tables + camera + lights + atlas, and the three corner-camera scenes
of tests/test_gpu_zz_edge_camera.py with their Oracle-B frames."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "fuzz"):
    sys.path.insert(0, str(p))
import bench  # noqa: E402
import fuzz_closed_form as F  # noqa: E402

O, pkg = F.O, F.pkg
S = pkg.scene
out = ROOT / "build" / "quick"
out.mkdir(parents=True, exist_ok=True)



def lights8(l):
    """eight light slots of ten floats (the layout vr_assign_lights is given), unused slots zero"""
    a = np.zeros((8, 10), np.float32)
    l = np.asarray(l, np.float32).reshape(-1, 10)
    a[: len(l)] = l
    return a


n = 1024
lo, hi = S.terrain_columns(n, "shell")
pos, direction = S.make_camera(n, S.heightfield(n), bench.BENCH_CAMERA)
lights = S.make_lights(n, 1)
sc = S.Scene(n, None, 3840, 2160, pos, direction, lights, max_distance=3 * n, columns=(lo, hi))
with open(out / "c3.bin", "wb") as f:
    np.array([n, 3840, 2160, 1, 3 * n], np.int32).tofile(f)
    np.asarray(pos, np.float32).tofile(f)
    np.asarray(direction, np.float32).tofile(f)
    lights8(lights).tofile(f)
    lo.astype(np.int32).tofile(f)
    hi.astype(np.int32).tofile(f)
np.ascontiguousarray(sc.atlas, np.uint8).tofile(out / "atlas.bin")
print("c3: atlas", sc.atlas.shape, "tile", sc.tile)

F.KINDS = "random,random,sparse,terrain,tunnel".split(",")
_, corner, nl, _ = F.make_case(1002, 4830)
_, biased, nl_b, _ = F.make_case(5002, 2093)
m = 64
p2, d2 = S.make_camera(m, S.heightfield(m), 2)
p2 = np.array([np.floor(p2[0]), np.floor(p2[1]), p2[2]], np.float32)
edge = S.Scene(m, S.terrain_map(m, "shell"), 160, 96, p2, d2, S.make_lights(m, 1), max_distance=20)
for name, scene, lights_n in (("corner", corner, nl), ("edge", edge, 1), ("biased", biased, nl_b)):
    desc, root = pkg.octree_generate(scene.volume)
    rgba, aux, _ = O.raycast(scene, octree=(desc, root), shadow_lights=lights_n, canonical_t=True)
    assert np.array_equal(scene.atlas, sc.atlas)
    with open(out / f"{name}.bin", "wb") as f:
        np.array([scene.n, scene.width, scene.height, lights_n, scene.max_distance], np.int32).tofile(f)
        np.asarray(scene.cam_pos, np.float32).tofile(f)
        np.asarray(scene.cam_dir, np.float32).tofile(f)
        lights8(scene.lights).tofile(f)
        np.ascontiguousarray(scene.volume, np.int8).tofile(f)
        np.ascontiguousarray(rgba, np.uint8).tofile(f)
    print(name, scene.n, scene.width, scene.height, lights_n, scene.max_distance, "written pixels differ from fill:", int((rgba != np.array([255, 255, 255, 100], np.uint8)).any(-1).sum()))
