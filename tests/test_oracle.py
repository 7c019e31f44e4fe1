"""CPU tests of the parity oracle itself (oracle/, a restatement of kernels/ray_caster_kernel.cl).

The reference ships no tests or golden vectors for this path ("parity unpinned"), so
the oracle is pinned by known answers DERIVED from the reference sources and by hand-computed rays."""
import numpy as np
import pytest

from conftest import oracle_bias


def test_head_octree_known_answer(pkg, oracle):
    """HEAD scene (16^3, all voxels 5, buffer of 100000, ref include/map/Octree.h:29): the generator
    writes 1+8+64+512 = 585 descriptors from index 99999 downward, so the root lands at 99415 with
    valid mask 0xFF, leaf mask 0 and relative pointer 1 (ref src/map/Octree.cpp:27-31)."""
    scene = pkg.scene.make_scene("head")
    buf, root, used = oracle.octree_generate(scene.volume, 100000)
    assert (root, used) == (99415, 585)
    assert int(buf[root]) == 0x00FF0001
    # the 8 children of the root sit right behind it, every 2^3-level descriptor is valid=leaf=0xFF
    assert all(((int(buf[root + 1 + i]) >> 16) & 0xFF) == 0xFF for i in range(8))
    assert int(buf[99999]) == 0xFFFF0000
    assert np.count_nonzero(buf) == 585


def test_octree_validate_property(pkg, oracle):
    """Octree::Validate (ref src/map/Octree.cpp:329-352): GetVoxel(pos).found == (dense voxel != 0)."""
    for name in ("tiny", "features"):
        scene = pkg.scene.make_scene(name)
        buf, root, used = oracle.octree_generate(scene.volume, 100000)
        assert root >= 0
        n = scene.n
        for z in range(n):
            for y in range(n):
                for x in range(0, n, 3):
                    found, sub, res, _ = oracle.get_oct_vox(buf, root, n, (x, y, z))
                    assert bool(found) == bool(scene.volume[z, y, x]), (x, y, z)
                    # the reported cell contains the query and is aligned to its own size
                    assert all(s <= p < s + max(res, 1) for s, p in zip(sub, (x, y, z)))


def test_octree_empty_collapse(pkg, oracle):
    """Only uniformly EMPTY subtrees collapse (ref src/map/Octree.cpp:230-233): a single voxel in a 16^3
    map gives root + 3 levels of one descriptor each; get_oct_vox far away stops at the root level."""
    vol = np.zeros((16, 16, 16), dtype=np.int8)
    vol[0, 0, 0] = 5
    buf, root, used = oracle.octree_generate(vol, 1000)
    assert used == 4
    found, sub, res, scale = oracle.get_oct_vox(buf, root, 16, (12, 3, 9))
    assert (found, sub, res, scale) == (0, (8, 0, 8), 8, 0)
    found, sub, res, scale = oracle.get_oct_vox(buf, root, 16, (0, 0, 0))
    assert (found, sub, res, scale) == (1, (0, 0, 0), 1, 3)
    found, sub, res, scale = oracle.get_oct_vox(buf, root, 16, (1, 0, 0))
    assert (found, sub, res, scale) == (0, (1, 0, 0), 1, 3)


def test_ray_table(oracle):
    """create_viewport (ref src/CLCaster.cpp:244-275): unit rays, focal length 800 px, centre pixel
    (x=0,y=0) is the base ray (-800,0,0) rotated by 1.57 rad about Y => ~(+0, 0, +1)."""
    w, h = 64, 36
    t = oracle.make_ray_table(w, h)
    assert t.shape == (h, w, 4) and np.all(t[..., 3] == 0)
    norms = np.sqrt((t[..., :3].astype(np.float64) ** 2).sum(-1))
    assert np.allclose(norms, 1.0, atol=1e-6)
    centre = t[h // 2, w // 2]
    assert abs(centre[2] - 1.0) < 1e-6 and centre[1] == 0.0 and abs(centre[0] + 800 * np.cos(1.57) / 800) < 1e-6
    # pixel x maps to ray.y, pixel y to ray.x (before the camera rotation): tan = offset / 800
    assert np.isclose(t[h // 2, w // 2 + 10, 1] / t[h // 2, w // 2 + 10, 2], 10 / 800.0, rtol=1e-3)
    # odd sizes leave the last row/column zero (loops run over 2*(size/2) entries)
    t2 = oracle.make_ray_table(5, 5)
    assert np.all(t2[4] == 0) and np.all(t2[:, 4] == 0) and np.all(t2[:4, :4, :3].any(-1))


def _single_pixel_scene(pkg, volume, pos, ray, light=(100.0, 100.0, 100.0), max_distance=64):
    """1x1 viewport whose only table entry is `ray` and an identity camera rotation (angles 0)."""
    S = pkg.scene
    n = volume.shape[0]
    lights = np.zeros((1, 10), np.float32)
    lights[0, :4] = (0.5, 0.5, 0.5, 1.0)
    lights[0, 4:7] = light
    scene = S.Scene(n, volume, 1, 1, np.array(pos, np.float32), np.array([0.0, 0.0], np.float32), lights,
                    max_distance=max_distance)
    table = np.zeros((1, 1, 4), np.float32)
    table[0, 0, :3] = ray
    return scene, table


def test_hand_computed_rays(pkg, oracle):
    """Known-answer rays in an 8^3 map worked out by hand from kernel:298-323,555-570.
    Camera (1.5, 1.25, 1.75), ray (0.8, 0.6, ~0) * small z: steps x,y alternate by t order."""
    vol = np.zeros((8, 8, 8), np.int8)
    vol[1, :, 6] = 5                    # wall x == 6 at z == 1
    d = np.array([0.8, 0.6, 1e-3], np.float64)
    d /= np.linalg.norm(d)
    scene, table = _single_pixel_scene(pkg, vol, (1.5, 1.25, 1.75), d.astype(np.float32))
    rgba, aux, cnt = oracle.raycast(scene, table, want_counters=True)
    a = aux[0, 0]
    # x crossings at t = 0.625 + 1.25k, y crossings at t = 1.25 + 1.667k: order x y x x y x => voxel (6, 4, 1)
    assert tuple(a["hit"]) == (6, 4, 1)
    assert a["face"] == 0b000001          # hit through the x face, all steps positive
    assert a["steps_first"] == 7           # 5 x-steps + 3 y-steps - 1 (distance counts completed iterations)
    assert a["flags"] & oracle.FL_LIT
    # a ray leaving the map without hitting anything writes (0,0,0,0)
    scene, table = _single_pixel_scene(pkg, vol, (1.5, 1.25, 5.5), np.array([0.6, 0.0, 0.8], np.float32) + np.float32(1e-4))
    rgba, aux, _ = oracle.raycast(scene, table)
    assert aux[0, 0]["status"] == oracle.ST_OOB and tuple(rgba[0, 0]) == (0, 0, 0, 0)
    # a zero component skips the pixel: it keeps the viewport's initial fill (255,255,255,100)
    scene, table = _single_pixel_scene(pkg, vol, (1.5, 1.25, 1.75), np.array([1.0, 0.0, 0.0], np.float32))
    rgba, aux, _ = oracle.raycast(scene, table)
    assert aux[0, 0]["status"] == oracle.ST_SKIP_PRIMARY and tuple(rgba[0, 0]) == (255, 255, 255, 100)


def test_max_distance_and_first_voxel_not_tested(pkg, oracle):
    """kernel:326,357: HEAD stops after 20 steps; kernel:555-570 steps BEFORE loading, so the camera's own
    voxel is never tested."""
    vol = np.zeros((64, 64, 64), np.int8)
    vol[10, 10, 40] = 5
    vol[10, 10, 5] = 5                   # the camera voxel itself
    ray = np.array([1.0, 1e-4, 1e-4], np.float32)
    scene, table = _single_pixel_scene(pkg, vol, (5.5, 10.5, 10.5), ray, max_distance=20)
    _, aux, _ = oracle.raycast(scene, table)
    assert aux[0, 0]["status"] == oracle.ST_MAXDIST and aux[0, 0]["steps_total"] == 20 and aux[0, 0]["hit"][0] == -1
    scene.max_distance = 64
    _, aux, _ = oracle.raycast(scene, table)
    assert tuple(aux[0, 0]["hit"]) == (40, 10, 10) and aux[0, 0]["steps_first"] == 34


def test_shadow_and_reflection_paths(pkg, oracle):
    """Outcome classes of SURVEY appendix A: lit (alpha = brightness), shadowed (alpha 0.1 -> 26),
    lit-but-shadow-ray-left-the-map (alpha 0), reflection (type 6) bounces limited to 2."""
    scene = pkg.scene.make_scene("features")
    rgba, aux, cnt = oracle.raycast(scene, want_counters=True)      # no octree bound: start bias 0
    lit = (aux["flags"] & oracle.FL_LIT) != 0
    shadowed = lit & (aux["status"] == oracle.ST_SHADOW_HIT)
    assert shadowed.sum() > 50 and cnt["reflect_rays"] > 100 and cnt["shadow_rays"] == lit.sum()
    # kernel:708 color.w = 0.1, fog factor (1 - steps/700) keeps it at round(0.1*255*f) <= 26
    assert rgba[shadowed][:, 3].max() <= 26 and rgba[shadowed][:, 3].min() >= 20
    assert np.all(aux["status"][(aux["flags"] & oracle.FL_REFLECTED) != 0] != oracle.ST_SKIP_PRIMARY)
    mirror = pkg.scene.make_scene("features-mirror")
    rgba, aux, cnt = oracle.raycast(mirror, want_counters=True)
    assert np.all(aux["status"] == oracle.ST_BOUNCES) and cnt["reflect_rays"] == 2 * cnt["primary_rays"]


def test_bias_changes_the_image(pkg, oracle):
    """kernel:353-354: the get_oct_vox result of the camera voxel biases intersection_t in BOTH modes; a
    camera in a collapsed empty cell therefore renders differently with and without an octree bound."""
    scene = pkg.scene.make_scene("features-high")
    desc, root = pkg.octree_generate(scene.volume)
    assert oracle_bias(oracle, scene, desc, root) == [-20, -12, -12]
    with_oct, _, _ = oracle.raycast(scene, octree=(desc, root))
    without, _, _ = oracle.raycast(scene)
    assert not np.array_equal(with_oct, without)


def test_thread_count_independence(pkg, oracle):
    scene = pkg.scene.make_scene("features-low")
    a, aa, _ = oracle.raycast(scene, threads=1)
    b, ab, _ = oracle.raycast(scene, threads=4)
    assert np.array_equal(a, b) and np.array_equal(aa, ab)


def test_golden_fixture(pkg, oracle):
    """Regression pin: frames generated by this oracle (tests/golden/make_golden.py) must not drift."""
    import pathlib

    g = pathlib.Path(__file__).parent / "golden"
    files = sorted(g.glob("*.npz"))
    assert files, "golden fixtures missing"
    for f in files:
        z = np.load(f)
        scene = pkg.scene.make_scene(str(z["scene"]))
        lights = int(z["lights"]) if "lights" in z.files else 1
        if lights > 1:
            from test_emu_parity import _with_lights

            scene = _with_lights(pkg, scene, lights)
        desc, root = pkg.octree_generate(scene.volume)
        rgba, aux, _ = oracle.raycast(scene, octree=(desc, root), shadow_lights=lights)
        assert np.array_equal(rgba, z["rgba"]), f.name
        for k in ("hit", "face", "status", "flags", "steps_first", "steps_total"):
            assert np.array_equal(aux[k], z[k]), (f.name, k)


def _golden_ref_scene(pkg, name):
    S = pkg.scene
    if name == "terrain64-cam3":
        n = 64
        vol = S.terrain_map(n, "shell", reflect_fraction=0.05)
        pos, direction = S.make_camera(n, S.heightfield(n), 3)
        return S.Scene(n, vol, 256, 144, pos, direction, S.make_lights(n), max_distance=3 * n)
    return S.make_scene(name)


def test_reference_generated_golden_vectors(pkg, oracle):
    """tests/golden/ref/*.npz were written by the REFERENCE'S OWN kernel and Octree::Generate executed on the CPU
    (tests/golden/make_golden_ref.py over oracle/_ref, see tests/test_reference_kernel.py): the oracle must reproduce
    every pixel, the written mask, the root index and every used descriptor -- a pin that needs nothing but git."""
    import pathlib

    g = pathlib.Path(__file__).parent / "golden" / "ref"
    frames = sorted(f for f in g.glob("*.npz") if not f.name.endswith("-octree.npz") and not f.name.startswith("viewport-"))
    assert len(frames) >= 16
    for f in frames:
        z = np.load(f)
        scene = _golden_ref_scene(pkg, str(z["scene"]))
        buf, root, _ = oracle.octree_generate(scene.volume, 100000)
        rgba, aux, _ = oracle.raycast(scene, octree=(buf, root), max_distance=int(z["max_distance"]))
        assert np.array_equal(rgba, z["rgba"]), f.name
        skipped = (aux["status"] == 0) | (aux["status"] == 4)
        assert np.array_equal(z["written"], ~skipped), f.name
    trees = sorted(g.glob("*-octree.npz"))
    assert len(trees) >= 3
    for f in trees:
        z = np.load(f)
        scene = _golden_ref_scene(pkg, str(z["scene"]))
        buf, root, _ = oracle.octree_generate(scene.volume, 100000)
        first = int(z["first_used"])
        assert root == int(z["root_index"]) and not buf[:first].any(), f.name
        assert np.array_equal(buf[first:], z["descriptors"]), f.name


def test_ray_table_equals_reference_generated_viewport_tables(oracle):
    """tests/golden/ref/viewport-tables.npz holds ray tables written by the loop of the reference's CLCaster::create_viewport
    (src/CLCaster.cpp:244-275, compiled from the reference's source: oracle/ref_shim/ref_viewport_host.cpp, generated by
    tests/golden/make_golden_ref.py): the oracle's table equals them bit for bit -- from git alone."""
    import pathlib

    z = np.load(pathlib.Path(__file__).parent / "golden" / "ref" / "viewport-tables.npz")
    for key, (w, h) in (("t64x36", (64, 36)), ("t5x7", (5, 7))):
        assert np.array_equal(oracle.make_ray_table(w, h).view(np.uint32), z[key].view(np.uint32)), key
    t = oracle.make_ray_table(3840, 2160)[::60, ::40]
    assert np.array_equal(t.view(np.uint32), z["t3840x2160_every_60th_row_40th_col"].view(np.uint32))


def test_descriptor_counters_from_a_4x4x4_occupancy_tree(pkg, oracle):
    """The D_svo byte model of 4096^3 (no dense map, hence no reference-format buffer) reads the path of 2^3 descriptors
    off an occupancy tree with 4^3 children per node (vro_scene::tree64): on maps that have both, the counters are the
    same as from the reference-format octree."""
    import emu_lib

    S = pkg.scene
    n = 64
    vol = S.terrain_map(n, "shell")
    pos, d = S.make_camera(n, S.heightfield(n), 3)
    scene = S.Scene(n, vol, 320, 180, pos, d, S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    _, _, a = oracle.raycast(scene, octree=(desc, root), want_aux=False, want_counters=True, count_svo=True)
    emu_lib.set_collapse(False)
    try:
        nodes, _, levels = emu_lib.tree_from_dense(vol)
    finally:
        emu_lib.set_collapse(True)
    _, _, b = oracle.raycast(scene, want_aux=False, want_counters=True, count_svo=True, tree64=(nodes, levels))
    assert a["svo_desc_fetches"] == b["svo_desc_fetches"] > 0 and a["svo_cell_changes"] == b["svo_cell_changes"]
