"""Long-running differential fuzz of the library's DEFAULT walk on the CPU: the closed-form walk (csrc/vr_canon.h, host build
tests/host_emu) over the undirected top grid (use_svo 3) and the directed ones (use_svo 4), with and without solid-subtree
collapse, against Oracle-B (the oracle with closed-form crossing times); bar = tests/test_gpu_canonical.assert_equals_oracle_b.

    python tests/fuzz/fuzz_closed_form.py SEED N

Scene kinds, drawn at random: the small random scenes of the suite (8^3..64^3, cameras inside / outside / on integer
coordinates, 1-3 lights, max_distance 5 / 20 / 3N), sparse 128^3 / 256^3 maps (cameras in collapsed empty cells: negative
start bias, cells hundreds of voxels wide), terrain 64^3..256^3 ("shell" and "solid", mirrors, 1 % holes, cameras on the ground
and high above it), solid blocks with carved tunnels (long runs of bricks next to collapsed solid nodes).
"""
import sys
import pathlib

R_ = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(R_ / "tests"))
sys.path.insert(0, str(R_))
import importlib

import numpy as np

import emu_lib
import oracle_lib as O
from conftest import oracle_bias
from test_emu_parity import random_scene
from test_gpu_canonical import assert_equals_oracle_b, sparse_scene

pkg = importlib.import_module("voxel-raycaster_b200")
S = pkg.scene
seed, count = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)


def terrain_scene():
    n = int(rng.choice([64, 128, 256]))
    kind = str(rng.choice(["shell", "solid"]))
    vol = S.terrain_map(n, kind, reflect_fraction=float(rng.choice([0.0, 0.002, 0.05]))).copy()
    if rng.random() < 0.4:
        holes = rng.random(vol.shape) < 0.01
        vol[holes & (vol != 0)] = 0
    pos, d = S.make_camera(n, S.heightfield(n), int(rng.integers(0, 40)))
    pos = np.array(pos, np.float32)
    d = np.array(d, np.float32)
    if rng.random() < 0.3:
        pos[2] = min(n - 1.5, pos[2] + float(rng.random() * n / 2))          # high above the ground
    nl = int(rng.choice([1, 2]))
    return S.Scene(n, vol, 160, 96, pos, d, S.make_lights(n, nl), max_distance=3 * n), nl


def tunnel_scene():
    n = int(rng.choice([32, 64, 128]))
    vol = np.full((n, n, n), 5, np.int8)
    vol[rng.random(vol.shape) < 0.01] = 6
    for _ in range(int(rng.integers(1, 6))):                                  # axis-aligned tunnels of random cross-section
        a = int(rng.integers(0, 3))
        lo = rng.integers(1, n - 4, size=3)
        w = rng.integers(1, 5, size=3)
        sl = [slice(int(lo[i]), int(lo[i] + w[i])) for i in range(3)]
        sl[a] = slice(0, n)
        vol[tuple(sl)] = 0
    empty = np.argwhere(vol == 0)
    z, y, x = empty[int(rng.integers(0, len(empty)))]
    pos = np.array([x + rng.random(), y + rng.random(), z + rng.random()], np.float32)
    d = np.array([rng.random() * np.pi, rng.random() * 2 * np.pi], np.float32)
    nl = int(rng.choice([1, 2]))
    lights = np.zeros((8, 10), np.float32)
    for l in range(nl):
        lz, ly, lx = empty[int(rng.integers(0, len(empty)))]
        lights[l] = [0.6, 0.6, 0.6, 1.0, lx + 0.5, ly + 0.5, lz + 0.5, -1, -1, -1.5]
    return S.Scene(n, vol, 96, 64, pos, d, lights, max_distance=3 * n), nl


bad = 0
ties = pixels = 0
for it in range(count):
    kind = str(rng.choice(["random", "random", "sparse", "terrain", "tunnel"]))
    if kind == "random":
        scene, nl = random_scene(pkg, rng)
        if scene.n < 8:
            continue
    elif kind == "sparse":
        scene, nl = sparse_scene(pkg, rng, int(rng.choice([128, 256])))
        scene.width, scene.height = 96, 64
    elif kind == "terrain":
        scene, nl = terrain_scene()
    else:
        scene, nl = tunnel_scene()
    table = O.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = O.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
    bias = oracle_bias(O, scene, desc, root)
    collapse = bool(rng.random() < 0.7)
    emu_lib.set_collapse(collapse)
    frames = []
    for use_svo in (3, 4):
        rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
        frames.append(rgba)
        try:
            ties += assert_equals_oracle_b(ref_rgba, ref_aux, rgba, aux, f"it {it} {kind} n={scene.n} svo={use_svo} collapse={collapse}")
            pixels += rgba.shape[0] * rgba.shape[1]
        except AssertionError as e:
            bad += 1
            print("MISMATCH", seed, it, kind, scene.n, use_svo, collapse, list(scene.cam_pos), list(scene.cam_dir), nl, scene.max_distance, str(e)[:160], flush=True)
    if not np.array_equal(frames[0], frames[1]):
        bad += 1
        print("GRIDS DIFFER", seed, it, kind, scene.n, flush=True)
    if it % 25 == 0:
        print(seed, it, kind, scene.n, "max steps", int(ref_aux["steps_total"].max()), "tie pixels so far", ties, "of", pixels, flush=True)
emu_lib.set_collapse(True)
print("done seed", seed, "scenes", count, "bad", bad, "tie pixels", ties, "of", pixels)
