"""CPU tests: the device core (voxel-raycaster_b200/csrc/vr_trace.h) compiled for the host must agree
bit for bit with the oracle -- RGBA8 and every integer aux field, on ALL pixels, for both the dense
DDA variant and the 64-tree variant.  This is a debugging aid for the control flow (a gpurun round trip
takes minutes); the GPU parity tests proper are in test_gpu_parity.py."""
import numpy as np
import pytest

import emu_lib
from conftest import assert_same_frame, oracle_bias

SCENES = ["head", "tiny", "small", "features", "features-low", "features-high", "features-mirror"]


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("use_svo", [False, True])
def test_device_core_matches_oracle(pkg, oracle, name, use_svo):
    scene = pkg.scene.make_scene(name)
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root))
    bias = oracle_bias(oracle, scene, desc, root)
    rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo)
    assert_same_frame(ref_rgba, ref_aux, rgba, aux, f"{name} svo={use_svo}")


@pytest.mark.parametrize("cam", [1, 3, 4])
def test_terrain_64(pkg, oracle, cam):
    S = pkg.scene
    n = 64
    vol = S.terrain_map(n, "shell", reflect_fraction=0.05)
    pos, direction = S.make_camera(n, S.heightfield(n), cam)
    scene = S.Scene(n, vol, 256, 144, pos, direction, S.make_lights(n), max_distance=3 * n)
    table = oracle.make_ray_table(scene.width, scene.height)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table)
    for use_svo in (False, True):
        rgba, aux = emu_lib.raycast(scene, table, use_svo=use_svo)
        assert_same_frame(ref_rgba, ref_aux, rgba, aux, f"terrain cam={cam} svo={use_svo}")
    assert (ref_aux["flags"] & 1).mean() > 0.2


def test_native_tree_matches_dense(pkg, oracle):
    """64-tree occupancy == (voxel in {5,6}) -- the Octree::Validate property for the native layout,
    checked through the traversal itself: with max_distance huge every SVO ray equals the dense ray."""
    S = pkg.scene
    rng = np.random.default_rng(3)
    n = 16
    vol = np.zeros((n, n, n), np.int8)
    vol[rng.random((n, n, n)) < 0.03] = 5
    vol[rng.random((n, n, n)) < 0.01] = 6
    vol[rng.random((n, n, n)) < 0.02] = 3     # transparent
    vol[8, 8, 8] = 0
    scene = S.Scene(n, vol, 96, 64, np.array([8.4, 8.6, 8.3], np.float32), np.array([1.9, 0.7], np.float32),
                    S.make_lights(n), max_distance=3 * n)
    table = oracle.make_ray_table(scene.width, scene.height)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table)
    rgba, aux = emu_lib.raycast(scene, table, use_svo=True)
    assert_same_frame(ref_rgba, ref_aux, rgba, aux, "random16")


@pytest.mark.parametrize("n,variant", [(16, "shell"), (64, "shell"), (64, "solid"), (128, "shell"), (256, "shell")])
def test_column_builder_equals_dense_builder(pkg, n, variant):
    """vr_native_from_columns (no N^3 volume, needed for 4096^3) must emit exactly the arrays
    vr_native_from_dense emits for the materialised map: same nodes, same order, same leaf types."""
    S = pkg.scene
    lo, hi = S.terrain_columns(n, variant)
    vol = S.terrain_map(n, variant)
    a_nodes, a_types, a_levels = emu_lib.tree_from_dense(vol)
    b_nodes, b_types, b_levels = emu_lib.tree_from_columns(lo, hi)
    assert a_levels == b_levels and a_nodes.shape == b_nodes.shape
    assert np.array_equal(a_nodes, b_nodes) and np.array_equal(a_types, b_types)


def test_column_builder_irregular(pkg):
    rng = np.random.default_rng(11)
    n = 32
    lo = rng.integers(0, n, size=(n, n)).astype(np.int32)
    hi = lo + rng.integers(-3, 6, size=(n, n)).astype(np.int32)       # some empty columns (hi < lo), some past the top
    vol = np.zeros((n, n, n), np.int8)
    for y in range(n):
        for x in range(n):
            if hi[y, x] >= lo[y, x]:
                vol[lo[y, x] : min(hi[y, x], n - 1) + 1, y, x] = 6
    a = emu_lib.tree_from_dense(vol)
    b = emu_lib.tree_from_columns(lo, hi, 6)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    empty = emu_lib.tree_from_columns(np.ones((16, 16), np.int32), np.zeros((16, 16), np.int32))
    assert empty[0].shape == (1, 4) and not empty[0][0, :2].any()


@pytest.mark.parametrize("name", SCENES)
def test_axis_walk_device_core(pkg, oracle, name):
    """Per-axis in-cell walk (option walk=1) compiled for the host: identical to the oracle on all non-tie
    pixels; tie pixels within +-1 RGBA (see vr_walk_axes)."""
    scene = pkg.scene.make_scene(name)
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root))
    rgba, aux = emu_lib.raycast(scene, table, bias=oracle_bias(oracle, scene, desc, root), use_svo=2)
    tie = (ref_aux["flags"] & 4) != 0
    for f in ("hit", "face", "status", "hit_type", "steps_first", "steps_total"):
        assert not (np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1) & ~tie).any(), f
    diff = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(-1)
    assert not (diff[~tie] > 0).any() and (diff <= 1).all()


def test_add_chain_binade_jumps():
    """vr_add_chain evaluates n literal float additions (kernel:559) in closed form per binade: bit-identical
    to the literal chain for random states, for d exactly half way between two grid points of t (ties-to-even
    alternation), for d below half an ulp of t (t stops moving), from t = 0 and from negative t."""
    rng = np.random.default_rng(1)
    n_s = 200000
    t = (rng.random(n_s) * np.exp2(rng.integers(-30, 13, n_s))).astype(np.float32)
    d = (1.0 / np.maximum(rng.random(n_s), 1e-6)).astype(np.float32)
    n = rng.integers(0, 3000, n_s).astype(np.int32)
    q = n_s // 4
    d[:q] = (np.round(d[:q] * 8) / 8 + np.exp2(-rng.integers(8, 20, q).astype(np.float32))).astype(np.float32)
    d[q:2 * q] = np.exp2(-rng.integers(0, 30, q).astype(np.float32))
    t[:1000] = 0
    a, b = emu_lib.add_chain(t, d, n)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    m = 20000
    e = rng.integers(1, 12, m)
    t2 = (np.exp2(e) * (1 + rng.random(m))).astype(np.float32)
    u = np.exp2(e - 23.0)
    d2 = (np.floor((1 + rng.random(m) * 3) / u) * u + u / 2).astype(np.float32)
    a, b = emu_lib.add_chain(t2, d2, rng.integers(8, 2000, m).astype(np.int32))
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    # negative start (the get_oct_vox bias of a camera in a collapsed empty cell, kernel:353): |t| shrinks first
    t3 = (-rng.random(m) * np.exp2(rng.integers(-3, 10, m))).astype(np.float32)
    d3 = (1.0 / np.maximum(rng.random(m), 1e-3)).astype(np.float32)
    a, b = emu_lib.add_chain(t3, d3, rng.integers(0, 1500, m).astype(np.int32))
    assert np.array_equal(a.view(np.int32), b.view(np.int32))


def test_axis_walk_terrain_256(pkg, oracle):
    """Per-axis walk with long chains (cells of up to 64^3 voxels, hundreds of crossings per cell) vs the oracle."""
    S = pkg.scene
    n = 256
    vol = S.terrain_map(n, "shell")
    for cam in (3, 7, 10):
        pos, direction = S.make_camera(n, S.heightfield(n), cam)
        scene = S.Scene(n, vol, 480, 270, pos, direction, S.make_lights(n), max_distance=3 * n)
        table = oracle.make_ray_table(scene.width, scene.height)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, table)
        rgba, aux = emu_lib.raycast(scene, table, use_svo=2)
        tie = (ref_aux["flags"] & 4) != 0
        for f in ("hit", "face", "status", "hit_type", "steps_first", "steps_total"):
            assert not (np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1) & ~tie).any(), f
        diff = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(-1)
        assert not (diff[~tie] > 0).any() and (diff <= 1).all()
        assert ref_aux["steps_total"].max() > 300


def _with_lights(pkg, scene, count):
    """The scene with `count` shadow lights (extension beyond the reference: see vr_next_light / the oracle)."""
    n = scene.n
    extra = np.array([[0.3, 0.5, 0.7, 1.0, 0.2 * n, 0.7 * n, 0.9 * n, -1.0, -1.0, -1.5],
                      [0.5, 0.2, 0.2, 1.0, 0.8 * n, 0.8 * n, 0.6 * n, -1.0, -1.0, -1.5]], np.float32)
    lights = np.zeros((8, 10), np.float32)
    lights[0] = scene.lights[0]
    lights[1:count] = extra[: count - 1]
    scene.lights = lights
    return scene


@pytest.mark.parametrize("name", ["features", "features-low", "features-mirror", "small"])
@pytest.mark.parametrize("count", [2, 3])
def test_multi_light_device_core(pkg, oracle, name, count):
    """Multi-light extension (LIGHT_COUNT > 1): the device core equals the oracle's restatement of the same
    extension on all pixels, dense and octree; and with count 1 both are the reference path (other tests)."""
    scene = _with_lights(pkg, pkg.scene.make_scene(name), count)
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, counters = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=count, want_counters=True)
    one_rgba, _, _ = oracle.raycast(scene, table, octree=(desc, root))
    lit = int(((ref_aux["flags"] & 1) != 0).sum())
    assert counters["shadow_rays"] >= 1.5 * lit
    assert lit == 0 or (one_rgba != ref_rgba).any()
    bias = oracle_bias(oracle, scene, desc, root)
    for use_svo in (0, 1):
        rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=count)
        assert_same_frame(ref_rgba, ref_aux, rgba, aux, f"{name} lights={count} svo={use_svo}")
    rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=2, shadow_lights=count)
    tie = (ref_aux["flags"] & 4) != 0
    for f in ("hit", "face", "status", "hit_type", "steps_first", "steps_total"):
        assert not (np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1) & ~tie).any(), f
    diff = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(-1)
    assert not (diff[~tie] > 0).any() and (diff <= 1).all()


def random_scene(pkg, rng):
    """One random scene of the differential fuzz tests (also used on the GPU): 8^3..64^3 maps of random density with
    mirrors and transparent values, cameras at random / integer / half-integer coordinates and outside the map, random
    and axis-aligned view directions, 1-3 random lights, max_distance 5 / 20 / 3N.  Returns (scene, light count)."""
    S = pkg.scene
    n = int(rng.choice([8, 16, 32, 64]))
    vol = np.zeros((n, n, n), np.int8)
    dens = rng.choice([0.002, 0.01, 0.05, 0.2, 0.6])
    vol[rng.random((n, n, n)) < dens] = 5
    vol[rng.random((n, n, n)) < dens * 0.3] = 6
    vol[rng.random((n, n, n)) < 0.02] = int(rng.integers(-5, 9))
    mode = int(rng.integers(0, 5))
    pos = (rng.random(3) * n).astype(np.float32)
    if mode == 1:
        pos = np.floor(pos).astype(np.float32)
    if mode == 2:
        pos = (np.floor(pos) + 0.5).astype(np.float32)
    if mode == 3:
        pos = (rng.random(3) * n * 1.5 - 0.25 * n).astype(np.float32)
    pos = np.clip(pos, -3, n + 3).astype(np.float32)
    d = np.array([rng.random() * np.pi, rng.random() * 2 * np.pi], np.float32)
    if mode == 4:
        d = np.array([rng.choice([0, np.pi / 2, np.pi, 1.57]), rng.choice([0, np.pi / 2, np.pi, 3 * np.pi / 2])], np.float32)
    nl = int(rng.choice([1, 1, 2, 3]))
    lights = np.zeros((8, 10), np.float32)
    for l in range(nl):
        lights[l] = [rng.random(), rng.random(), rng.random(), rng.random() * 2, *(rng.random(3) * n * 1.2 - 0.1 * n), -1, -1, -1.5]
    return S.Scene(n, vol, 48, 32, pos, d, lights, max_distance=int(rng.choice([20, 3 * n, 5]))), nl


def assert_walk_matches(ref_rgba, ref_aux, rgba, aux, per_axis, what):
    """all pixels identical; for the per-axis walk the oracle-flagged tie pixels get the north-star tolerance"""
    tie = ((ref_aux["flags"] & 4) != 0) if per_axis else np.zeros(ref_aux["flags"].shape, bool)
    for f in ("hit", "face", "status", "hit_type", "steps_first", "steps_total"):
        assert not (np.any(np.atleast_3d(ref_aux[f] != aux[f]), axis=-1) & ~tie).any(), (what, f)
    diff = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(-1)
    assert not (diff[~tie] > 0).any() and (diff <= 1).all(), what


@pytest.mark.parametrize("seed", [0, 1])
def test_device_core_random_scenes(pkg, oracle, seed):
    """Differential fuzzing of the device core (host build) against the oracle: 30 random scenes per seed, dense /
    octree merged walk / per-axis walk, 1-3 lights.  (600 such scenes x 3 walks were run when this was introduced.)"""
    rng = np.random.default_rng(seed)
    for it in range(30):
        scene, nl = random_scene(pkg, rng)
        table = oracle.make_ray_table(scene.width, scene.height)
        desc, root = pkg.octree_generate(scene.volume)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl)
        bias = oracle_bias(oracle, scene, desc, root)
        for use_svo in (0, 1, 2):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_walk_matches(ref_rgba, ref_aux, rgba, aux, use_svo == 2, f"seed {seed} scene {it} svo {use_svo}")


def test_device_core_equals_reference_generated_golden_vectors(pkg, oracle):
    """The device core (host build), dense and octree, against the committed outputs of the reference's own kernel
    (tests/golden/ref): every pixel, and unwritten pixels keep create_viewport's initial fill."""
    import pathlib

    from test_oracle import _golden_ref_scene

    g = pathlib.Path(__file__).parent / "golden" / "ref"
    frames = sorted(f for f in g.glob("*.npz") if not f.name.endswith("-octree.npz") and not f.name.startswith("viewport-"))
    assert len(frames) >= 16
    for f in frames:
        z = np.load(f)
        scene = _golden_ref_scene(pkg, str(z["scene"]))
        scene.max_distance = int(z["max_distance"])
        desc, root = pkg.octree_generate(scene.volume)
        table = oracle.make_ray_table(scene.width, scene.height)
        bias = oracle_bias(oracle, scene, desc, root)
        for use_svo in (0, 1):
            rgba, _ = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo)
            assert np.array_equal(rgba, z["rgba"]), (f.name, use_svo)
            assert (rgba[~z["written"]] == np.array([255, 255, 255, 100], np.uint8)).all()


def test_device_core_random_sparse_maps(pkg, oracle):
    """Differential fuzzing where the octree cells are huge: 128^3 maps with 0.02-0.5 % random voxels (and sometimes a
    ground slab), cameras anywhere -- most sit in a collapsed empty cell, so the get_oct_vox bias makes intersection_t
    start NEGATIVE and the per-axis walk runs chains of hundreds of additions through zero (a case the binade jumps
    once got wrong).  Octree kernel core, merged and per-axis walk, 1-2 lights, against the oracle."""
    S = pkg.scene
    rng = np.random.default_rng(1)
    n = 128
    biased = 0
    for it in range(12):
        vol = np.zeros((n, n, n), np.int8)
        dens = rng.choice([0.0002, 0.001, 0.005])
        vol[rng.random((n, n, n)) < dens] = 5
        vol[rng.random((n, n, n)) < dens * 0.2] = 6
        if rng.random() < 0.5:
            vol[: n // 8] = 5
        pos = (rng.random(3) * n).astype(np.float32)
        pos[2] = max(pos[2], n // 8 + 1.3)
        d = np.array([rng.random() * np.pi, rng.random() * 2 * np.pi], np.float32)
        nl = int(rng.choice([1, 2]))
        lights = np.zeros((8, 10), np.float32)
        for l in range(nl):
            lights[l] = [rng.random(), rng.random(), rng.random(), 1.0, *(rng.random(3) * n), -1, -1, -1.5]
        scene = S.Scene(n, vol, 96, 64, pos, d, lights, max_distance=3 * n)
        table = oracle.make_ray_table(96, 64)
        desc, root = pkg.octree_generate(vol)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl)
        bias = oracle_bias(oracle, scene, desc, root)
        biased += any(b != 0 for b in bias)
        for use_svo in (1, 2):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_walk_matches(ref_rgba, ref_aux, rgba, aux, use_svo == 2, f"sparse scene {it} svo {use_svo}")
    assert biased >= 6


def test_reference_octree_importer_random_volumes(pkg):
    """vr_native_from_ref (what `assign_octree` without `assign_map` leads to): the 64-tree imported from a reference-
    format descriptor buffer equals the one built from the dense map (all solids typed 5) -- random volumes, including
    one whose buffer needs far pointers (> 32k descriptors, kernel:222-225) and a root wider than the map (32^3)."""
    rng = np.random.default_rng(9)
    for n, dens in ((8, 0.3), (16, 0.05), (32, 0.02), (32, 0.6), (64, 0.01), (128, 0.02)):
        vol = (rng.random((n, n, n)) < dens).astype(np.int8) * 5
        desc, root = pkg.octree_generate(vol)
        if n == 128:
            assert desc.size > 0x8000
        a = emu_lib.tree_from_dense(vol)
        b = emu_lib.tree_from_ref(desc, root, n)
        assert a[2] == b[2] and a[0].shape == b[0].shape
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (n, dens)


@pytest.mark.parametrize("kind", ["terrain64", "terrain256", "random32", "sparse128", "half32"])
def test_top_grid_invariants(pkg, kind):
    """The top grid of the closed-form walk (vr_octree.cpp: vr_native_grid): a block is marked non-empty exactly when it
    holds a set voxel, its node index is the node covering it, and every empty entry describes a cell that contains the
    block, lies inside the map and holds no set voxel (checked against the dense map)."""
    S = pkg.scene
    rng = np.random.default_rng(8)
    if kind.startswith("terrain"):
        vol = S.terrain_map(int(kind[7:]), "shell")
    elif kind == "random32":
        vol = (rng.random((32, 32, 32)) < 0.03).astype(np.int8) * 5
    elif kind == "sparse128":
        vol = (rng.random((128, 128, 128)) < 0.0003).astype(np.int8) * 6
    else:
        vol = np.zeros((32, 32, 32), np.int8)
        vol[:16] = 5
    n = vol.shape[0]
    nodes, types, levels = emu_lib.tree_from_dense(vol)
    grid, g, bits = emu_lib.grid_from_tree(nodes, levels, n)
    G = 1 << bits
    assert G == n >> g and grid.shape == (G, G, G)
    solid = (vol == 5) | (vol == 6)
    blocks = solid.reshape(G, 1 << g, G, 1 << g, G, 1 << g).any(axis=(1, 3, 5))          # [bz, by, bx]
    assert np.array_equal((grid & 0x80000000) != 0, blocks)
    # summed-area table of the set voxels: emptiness of a box in O(1)
    sat = np.zeros((n + 1, n + 1, n + 1), np.int64)
    sat[1:, 1:, 1:] = solid.astype(np.int64).cumsum(0).cumsum(1).cumsum(2)

    def count(lo, hi):          # set voxels in [lo, hi) per axis, (z, y, x)
        (z0, y0, x0), (z1, y1, x1) = lo, hi
        return (sat[z1, y1, x1] - sat[z0, y1, x1] - sat[z1, y0, x1] - sat[z1, y1, x0]
                + sat[z0, y0, x1] + sat[z0, y1, x0] + sat[z1, y0, x0] - sat[z0, y0, x0])

    wide = 0
    for bz, by, bx in np.argwhere(~blocks):
        e = int(grid[bz, by, bx])
        m, ext = (1 << (e & 31)) - 1, e >> 8
        lo, hi = [], []
        for b in (bz, by, bx):
            p = b << g
            lo.append((p & ~m) - ext)          # the cell is symmetric: it serves rays of either direction (mirroring)
            hi.append((p | m) + ext + 1)
        assert min(lo) >= 0 and max(hi) <= n, (kind, bz, by, bx, e)
        assert count(lo, hi) == 0, (kind, bz, by, bx, e)
        wide += hi[0] - lo[0] > (1 << g)
    if kind != "random32":
        assert wide > 0
