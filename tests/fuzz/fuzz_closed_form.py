"""Long-running differential fuzz of the library's DEFAULT walk on the CPU: the closed-form walk (csrc/vr_canon.h, host build
tests/host_emu) over the undirected top grid (use_svo 3) and the directed ones (use_svo 4), with and without solid-subtree
collapse, against Oracle-B (the oracle with closed-form crossing times); bar: classify() below.

    python tests/fuzz/fuzz_closed_form.py SEED N [FIRST]      (scene i of a seed depends on (SEED, i) only: FIRST replays from there)

Scene kinds, drawn at random: the small random scenes of the suite (8^3..64^3, cameras inside / outside / on integer
coordinates, 1-3 lights, max_distance 5 / 20 / 3N), sparse 128^3 / 256^3 maps (cameras in collapsed empty cells: negative
start bias, cells hundreds of voxels wide), terrain 64^3..256^3 ("shell" and "solid", mirrors, 1 % holes, cameras on the ground
and high above it), solid blocks with carved tunnels (long runs of bricks next to collapsed solid nodes).
"""
import os
import pathlib
import sys

R_ = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(R_ / "tests"))
sys.path.insert(0, str(R_))
import importlib

import numpy as np

import emu_lib
import oracle_lib as O
from conftest import oracle_bias
from test_emu_parity import assert_walk_matches, random_scene
from test_gpu_canonical import sparse_scene

pkg = importlib.import_module("voxel-raycaster_b200")
S = pkg.scene
_args = sys.argv[1:] if __name__ == "__main__" else []            # (imported by tests/test_emu_canonical.py for make_case / classify)
seed = int(_args[0]) if len(_args) > 0 else 0
count = int(_args[1]) if len(_args) > 1 else 0
first = int(_args[2]) if len(_args) > 2 else 0
rng = None                                   # set per scene (make_case)
LITERAL = os.environ.get("FUZZ_LITERAL", "0") == "1"        # also: dense / merged walk / per-axis walk against Oracle-A
KINDS = os.environ.get("FUZZ_KINDS", "random,random,sparse,terrain,tunnel").split(",")     # e.g. FUZZ_KINDS=tunnel


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


INT_FIELDS = ("hit", "face", "status", "hit_type", "steps_first", "steps_total")
HIT_FIELDS = ("hit", "face", "hit_type")


def make_case(seed, it):
    """scene number `it` of `seed` -> (kind, scene, light count, collapse); None for maps too small for a top grid"""
    global rng
    rng = np.random.default_rng([seed, it])
    kind = str(rng.choice(KINDS))
    if kind == "random":
        scene, nl = random_scene(pkg, rng)
        if scene.n < 8:
            return None
    elif kind == "sparse":
        scene, nl = sparse_scene(pkg, rng, int(rng.choice([128, 256])))
        scene.width, scene.height = 96, 64
    elif kind == "terrain":
        scene, nl = terrain_scene()
    else:
        scene, nl = tunnel_scene()
    return kind, scene, nl, bool(rng.random() < 0.7)


def classify(ref_rgba, ref_aux, rgba, aux, lights=1):
    """-> (problems, tie pixels, tie pixels beyond +-1 that the step count explains).
    Non-tie pixels: everything identical.  Tie pixels (Oracle-B saw an exact multi-axis step): same first hit; RGBA8 within
    +-1 -- or the walk counted an unobserved tie strictly inside an empty cell as two steps, its step
    count is then higher than Oracle-B's and a ray that ends by max_distance (kernel:357) ends one step earlier: the shadow
    ray of such a pixel may stop short of the voxel that Oracle-B's still reaches.  With several lights the final counter is
    that of the LAST light's ray, so an early end of an earlier light's ray cannot be read off it: such pixels (same first
    hit, same terminal status) are counted as explained too."""
    problems = []
    tie = (ref_aux["flags"] & 4) != 0
    for f in INT_FIELDS:
        b = np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1) & ~tie
        if b.any():
            problems.append(f"{f} differs on {int(b.sum())} non-tie pixels, first (y,x)={np.argwhere(b)[0].tolist()}")
    b = ((ref_aux["flags"] & 0xFB) != (aux["flags"] & 0xFB)) & ~tie
    if b.any():
        problems.append(f"flags differ on {int(b.sum())} non-tie pixels")
    diff = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(axis=-1)
    if (diff[~tie] > 0).any():
        problems.append(f"RGBA differs on {int((diff[~tie] > 0).sum())} non-tie pixels, first (y,x)={np.argwhere((diff > 0) & ~tie)[0].tolist()}")
    for f in HIT_FIELDS:
        b = np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1)
        if b.any():
            problems.append(f"first hit ({f}) differs on {int(b.sum())} tie pixels")
    far = tie & (diff > 1)
    explained = far & (aux["steps_total"].astype(np.int64) > ref_aux["steps_total"].astype(np.int64))
    if lights > 1:
        explained |= far & (aux["status"] == ref_aux["status"])
    if (far & ~explained).any():
        problems.append(f"{int((far & ~explained).sum())} tie pixels beyond +-1 without a higher step count, first (y,x)={np.argwhere(far & ~explained)[0].tolist()}")
    return problems, int(tie.sum()), int(explained.sum())


def main():
    bad = edge_frames = voxelwise_frames = ties = far = pixels = grids = scenes = 0
    for it in range(first, first + count):
        case = make_case(seed, it)
        if case is None:
            continue
        kind, scene, nl, collapse = case
        scenes += 1
        table = O.make_ray_table(scene.width, scene.height)
        desc, root = pkg.octree_generate(scene.volume)
        ref_rgba, ref_aux, _ = O.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
        bias = oracle_bias(O, scene, desc, root)
        on_edge = int((scene.cam_pos == np.floor(scene.cam_pos)).sum()) >= 2
        edge_frames += on_edge                # camera on a voxel edge / corner: vr_cam_on_edge (csrc/vr_types.h)
        voxelwise_frames += on_edge and any(int(b) != 0 for b in bias)
        emu_lib.set_collapse(collapse)
        frames = []
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            frames.append(rgba)
            problems, t, f = classify(ref_rgba, ref_aux, rgba, aux, nl)
            ties += t
            far += f
            pixels += rgba.shape[0] * rgba.shape[1]
            if f:
                print("tie pixel(s) beyond +-1, explained by the step count:", seed, it, kind, scene.n, use_svo, f, flush=True)
            if problems:
                bad += 1
                print("MISMATCH", seed, it, kind, scene.n, use_svo, collapse, nl, scene.max_distance, "; ".join(problems)[:300], flush=True)
        if LITERAL:
            # the bit-exact walks of round 1 over the same scene (walk 0: every pixel; walk 1: step count on tie pixels aside)
            a_rgba, a_aux, _ = O.raycast(scene, table, octree=(desc, root), shadow_lights=nl)
            for use_svo in (0, 1, 2):
                rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
                try:
                    assert_walk_matches(a_rgba, a_aux, rgba, aux, use_svo == 2, f"use_svo={use_svo}")
                except AssertionError as e:
                    bad += 1
                    print("MISMATCH (literal walk)", seed, it, kind, scene.n, use_svo, collapse, nl, str(e)[:200], flush=True)
        # the two kinds of top grid hand a ray different empty cells: the frames are equal except where a tie falls strictly
        # inside a cell of one grid and on a cell boundary of the other (tie pixels only)
        d = np.abs(frames[0].astype(np.int16) - frames[1].astype(np.int16)).max(axis=-1)
        tie = (ref_aux["flags"] & 4) != 0
        if (d[~tie] > 0).any():
            bad += 1
            print("GRIDS DIFFER on non-tie pixels", seed, it, kind, scene.n, flush=True)
        grids += int((d > 0).sum())
        if it % 100 == 0:
            print(seed, it, kind, scene.n, "max steps", int(ref_aux["steps_total"].max()), "tie pixels so far", ties, "of", pixels, flush=True)
    emu_lib.set_collapse(True)
    print(f"done seed {seed} scenes {scenes} (x 2 grids) unexplained {bad} edge-camera scenes {edge_frames} (voxelwise {voxelwise_frames}) pixels {pixels} tie pixels {ties} "
          f"tie pixels beyond +-1 (step count) {far} pixels on which the two grids differ (tie pixels) {grids}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
