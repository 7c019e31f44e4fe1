"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(libvrcaster.so via voxel-raycaster_b200.CUDACaster), against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): hit voxel index and face bit-exact, RGBA8 max abs diff <= 1 on >= 99.9 % of
pixels.  What is asserted here is stricter: RGBA8 and every integer aux field are IDENTICAL on ALL pixels,
for the dense DDA kernel and for the 64-tree kernel alike (tolerance 0, written in assert_same_frame)."""
import numpy as np
import pytest

from conftest import AUX_FIELDS, assert_same_frame, oracle_bias

pytestmark = pytest.mark.gpu

SMALL = ["head", "tiny", "small", "features", "features-low", "features-high", "features-mirror"]


def make_caster(pkg, scene, use_octree, assign_octree=True, aux=True, walk=0):
    """walk 0 (literal additions, merged walk: bit-identical to the reference) unless a test asks for another; the library
    default is walk 2 (test_gpu_canonical.py)"""
    c = pkg.CUDACaster()
    c.load_scene(scene, use_octree=use_octree, assign_octree=assign_octree, walk=walk)
    if aux:
        assert c.enable_aux(True)
    return c


def test_ray_table_bit_exact(pkg, oracle):
    """create_viewport (ref src/CLCaster.cpp:244-275): the device-resident table equals the restated one."""
    for w, h in ((50, 50), (64, 36), (1280, 720), (5, 7)):
        c = pkg.CUDACaster()
        assert c.init(0)
        assert c.create_viewport(w, h, 56.25, 90.0)
        assert np.array_equal(c.read_ray_table().view(np.uint32), oracle.make_ray_table(w, h).view(np.uint32))
        c.close()
    # and the tables the loop of the reference's own create_viewport wrote (tests/golden/ref/viewport-tables.npz)
    import pathlib

    z = np.load(pathlib.Path(__file__).parent / "golden" / "ref" / "viewport-tables.npz")
    for key, (w, h), sl in (("t64x36", (64, 36), (slice(None), slice(None))), ("t5x7", (5, 7), (slice(None), slice(None))),
                            ("t3840x2160_every_60th_row_40th_col", (3840, 2160), (slice(None, None, 60), slice(None, None, 40)))):
        c = pkg.CUDACaster()
        assert c.init(0) and c.create_viewport(w, h, 56.25, 90.0)
        assert np.array_equal(c.read_ray_table().reshape(h, w, 4)[sl].view(np.uint32), z[key].view(np.uint32)), key
        c.close()


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("use_octree", [False, True])
def test_small_scenes(pkg, oracle, name, use_octree):
    scene = pkg.scene.make_scene(name)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
    c = make_caster(pkg, scene, use_octree)
    assert c.compute(), c.last_error()
    st = c.stats()
    assert list(st.bias) == oracle_bias(oracle, scene, desc, root)
    assert st.used_svo == int(use_octree) and st.kernel_launches >= 1
    assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"{name} use_octree={use_octree}")
    c.close()


@pytest.mark.parametrize("name", ["head", "features", "features-low", "features-high", "features-mirror"])
def test_golden_fixtures(pkg, name):
    """The committed fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
    from pathlib import Path

    z = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
    scene = pkg.scene.make_scene(name)
    for use_octree in (False, True):
        c = make_caster(pkg, scene, use_octree)
        assert c.compute(), c.last_error()
        got, aux = c.draw(), c.read_aux()
        assert np.array_equal(got, z["rgba"])
        for k in ("hit", "face", "status", "flags", "hit_type", "steps_first", "steps_total"):
            assert np.array_equal(aux[k], z[k]), (name, k)
        c.close()


def test_head_defaults(pkg, oracle):
    """HEAD scene with HEAD's max_distance of 20 (kernel:326) when no MAX_DISTANCE setting is registered."""
    scene = pkg.scene.make_scene("head")
    c = pkg.CUDACaster()
    assert c.init(0)
    assert c.add_to_settings_buffer("octree_dimensions", "OCTDIM", 16)
    assert c.add_to_settings_buffer("using_octree", "OCTENABLED", 1)
    desc, root = pkg.octree_generate(scene.volume)
    assert c.assign_octree(desc, root) and c.assign_map(scene.volume)
    assert c.assign_camera(scene.cam_dir, scene.cam_pos) and c.create_viewport(50, 50, 56.25, 90.0)
    assert c.assign_lights(scene.lights) and c.create_texture_atlas(scene.atlas, (16, 16)) and c.validate()
    assert list(c.settings()[:3]) == [16, 1, root]          # slot order = call order (ref host:1047-1049)
    assert c.compute()
    ref, _, _ = oracle.raycast(scene, octree=(desc, root), max_distance=20)
    assert np.array_equal(c.draw(), ref)
    c.close()


@pytest.mark.parametrize("cam", [0, 1, 3, 4, 7])
def test_c1_terrain_64(pkg, oracle, cam):
    """BASELINE config 1 size: 64^3 procedural map, 1280x720, primary + 1 shadow light (+5 % mirrors)."""
    S = pkg.scene
    n = 64
    vol = S.terrain_map(n, "shell", reflect_fraction=0.05 if cam % 2 else 0.0)
    pos, direction = S.make_camera(n, S.heightfield(n), cam)
    scene = S.Scene(n, vol, 1280, 720, pos, direction, S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
    for use_octree in (False, True):
        c = make_caster(pkg, scene, use_octree)
        assert c.compute(), c.last_error()
        assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"c1 cam={cam} use_octree={use_octree}")
        c.close()


@pytest.mark.parametrize("variant", ["shell", "solid"])
def test_c2_terrain_256(pkg, oracle, variant):
    """BASELINE config 2 size: 256^3 terrain, 1920x1080, SVO kernel, shading + atlas + shadow light."""
    S = pkg.scene
    n = 256
    vol = S.terrain_map(n, variant)
    pos, direction = S.make_camera(n, S.heightfield(n), 4)
    scene = S.Scene(n, vol, 1920, 1080, pos, direction, S.make_lights(n), max_distance=3 * n)
    ref_rgba, ref_aux, _ = oracle.raycast(scene)
    for use_octree in (False, True):
        c = make_caster(pkg, scene, use_octree, assign_octree=False)
        assert c.compute(), c.last_error()
        assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"c2 {variant} use_octree={use_octree}")
        c.close()


def test_octree_only_import(pkg, oracle):
    """assign_octree without assign_map: the 64-tree is imported from the reference-format descriptors."""
    scene = pkg.scene.make_scene("small")
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
    c = pkg.CUDACaster()
    assert c.init(0)
    assert c.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n)
    assert c.add_to_settings_buffer("using_octree", "OCTENABLED", 0)
    assert c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance)
    assert c.assign_octree(desc, root)
    assert c.assign_camera(scene.cam_dir, scene.cam_pos) and c.create_viewport(scene.width, scene.height)
    assert c.assign_lights(scene.lights) and c.create_texture_atlas(scene.atlas) and c.validate(), c.last_error()
    assert c.set_option("walk", 0)
    assert c.enable_aux(True) and c.compute(), c.last_error()
    assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), "octree-only")
    # dense traversal without a map must fail loudly, not fall back
    assert c.overwrite_setting("using_octree", 1)
    assert not c.compute() and "no map" in c.last_error()
    c.close()


def test_host_pointer_aliasing(pkg, oracle):
    """CL_MEM_USE_HOST_PTR semantics (ref src/CLCaster.cpp:137-139,322,1069): camera, lights and settings are
    re-read from the caller's memory at every compute()."""
    scene = pkg.scene.make_scene("features")
    desc, root = pkg.octree_generate(scene.volume)
    c = make_caster(pkg, scene, True)
    assert c.compute()
    first = c.draw()
    scene.cam_pos[:] = (20.4, 5.6, 9.35)          # in place: same arrays the caster aliases
    scene.cam_dir[:] = (1.65, 1.9)
    scene.lights[0, 4:7] = (8.0, 20.0, 18.5)
    assert c.compute()
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
    got = c.draw()
    assert not np.array_equal(first, got)
    assert_same_frame(ref_rgba, ref_aux, got, c.read_aux(), "aliased update")
    c.settings()[2] = 7                             # MAX_DISTANCE slot, written through the aliased array
    assert c.compute()
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root), max_distance=7)
    assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), "aliased setting")
    c.close()


def test_bands_and_pipelined_frames(pkg, oracle):
    """Multi-GPU screen-tile split on one device: the interleaved band slabs reassemble to the full frame
    bit for bit; frame_begin/frame_end (async D2H, double buffered) return the same pixels as compute()."""
    scene = pkg.scene.make_scene("features-low")
    ref_rgba, ref_aux, _ = oracle.raycast(scene)
    H, W = scene.height, scene.width
    for band_rows, stride in ((8, 2), (8, 4), (16, 3), (7, 2)):
        out = np.zeros((H, W, 4), np.uint8)
        nb = (H + band_rows - 1) // band_rows
        for first in range(stride):
            c = pkg.CUDACaster()
            c.load_scene(scene, use_octree=True, assign_octree=False, walk=0)
            assert c.set_bands(band_rows, stride, first)
            assert c.compute(), c.last_error()
            slab = c.draw()
            rows = [r for b in range(first, nb, stride) for r in range(b * band_rows, min(H, (b + 1) * band_rows))]
            assert slab.shape[0] == len(rows) == c.local_rows()
            out[rows] = slab
            c.close()
        assert np.array_equal(out, ref_rgba), (band_rows, stride)
    c = pkg.CUDACaster()
    c.load_scene(scene, use_octree=True, assign_octree=False, walk=0)
    assert c.frame_begin() and c.frame_begin() and not c.frame_begin()      # at most two frames in flight
    a = c.frame_end().copy()
    b = c.frame_end().copy()
    assert np.array_equal(a, ref_rgba) and np.array_equal(b, ref_rgba)
    c.close()


def test_native_tree_broadcast_roundtrip(pkg, oracle):
    """The multi-GPU scene replication path on one device: export the 64-tree, adopt it in a second context
    that never saw the map, render the same frame."""
    import torch

    scene = pkg.scene.make_scene("features")
    a = pkg.CUDACaster()
    a.load_scene(scene, use_octree=True, assign_octree=False, walk=0)
    nb, tb, levels, dim = a.native_tree_info()
    nodes = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
    types = torch.empty(tb, dtype=torch.uint8, device="cuda:0")
    assert a.native_tree_copy(nodes.data_ptr(), types.data_ptr())
    assert a.compute()
    want = a.draw()
    b = pkg.CUDACaster()
    assert b.init(0)
    assert b.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n)
    assert b.add_to_settings_buffer("using_octree", "OCTENABLED", 0)
    assert b.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance)
    assert b.assign_native_tree(nodes.data_ptr(), nb, types.data_ptr(), tb, levels, dim), b.last_error()
    assert b.assign_camera(scene.cam_dir, scene.cam_pos) and b.create_viewport(scene.width, scene.height)
    assert b.assign_lights(scene.lights) and b.create_texture_atlas(scene.atlas) and b.validate(), b.last_error()
    assert b.set_option("walk", 0) and b.compute(), b.last_error()
    assert np.array_equal(b.draw(), want)
    a.close()
    b.close()


def test_error_behaviour(pkg):
    """CLCaster returns false and logs instead of throwing (ref src/CLCaster.cpp:157-185, 1029-1109)."""
    c = pkg.CUDACaster()
    assert c.init(0)
    assert not c.validate() and "camera" in c.last_error()
    assert c.add_to_settings_buffer("a", "A", 1) and not c.add_to_settings_buffer("b", "A", 2)     # duplicate define
    for i in range(63):
        assert c.add_to_settings_buffer(f"s{i}", f"S{i}", i)
    assert not c.add_to_settings_buffer("overflow", "OVERFLOW", 0) and "maximum size" in c.last_error()
    assert not c.overwrite_setting("nope", 3) and not c.remove_from_settings_buffer("a")
    assert not c.release_map() and not c.release_viewport()
    assert not c.compute()
    c.close()


def test_full_size_c3_properties(pkg, oracle):
    """BASELINE full size (1024^3 SVO, 3840x2160): size-independent properties.
    (1) 64-tree kernel frame == dense DDA kernel frame, bit for bit (both CUDA, 8.3 M pixels);
    (2) every 90th row equals the oracle; (3) determinism: two SVO frames are identical."""
    import bench

    scene = bench.bench_scene("c3")
    frames = {}
    for use_octree in (True, False):
        c = make_caster(pkg, scene, use_octree, assign_octree=False)
        assert c.compute(), c.last_error()
        frames[use_octree] = (c.draw(), c.read_aux())
        if use_octree:
            assert c.compute()
            assert np.array_equal(c.draw(), frames[True][0])
        c.close()
    assert np.array_equal(frames[True][0], frames[False][0])
    for f in AUX_FIELDS:
        assert np.array_equal(frames[True][1][f], frames[False][1][f]), f
    ref_rgba, ref_aux, _ = oracle.raycast(scene, row_stride=90)
    assert_same_frame(ref_rgba[::90], ref_aux[::90], frames[True][0][::90], frames[True][1][::90], "c3 rows")


@pytest.mark.parametrize("refill_min", [1, 8, 24])
def test_persistent_warps(pkg, oracle, refill_min):
    """Persistent-warp variant of the octree kernel (warp-level pixel refill through __ballot_sync/__shfl_sync):
    same pixels and aux records as the oracle, whatever the refill threshold."""
    for name in ("features", "features-low", "small"):
        scene = pkg.scene.make_scene(name)
        desc, root = pkg.octree_generate(scene.volume)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
        c = make_caster(pkg, scene, True)
        assert c.set_option("persistent", 1) and c.set_option("refill_min", refill_min)
        assert c.compute(), c.last_error()
        assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"persistent {name} refill={refill_min}")
        c.close()
    S = pkg.scene
    vol = S.terrain_map(64, "shell", reflect_fraction=0.05)
    pos, direction = S.make_camera(64, S.heightfield(64), 3)
    scene = S.Scene(64, vol, 1280, 720, pos, direction, S.make_lights(64), max_distance=192)
    ref_rgba, ref_aux, _ = oracle.raycast(scene)
    c = make_caster(pkg, scene, True, assign_octree=False)
    assert c.set_option("persistent", 1) and c.set_option("refill_min", refill_min)
    assert c.compute(), c.last_error()
    assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"persistent c1 refill={refill_min}")
    c.close()


def test_columns_equal_dense_1024(pkg):
    """assign_columns (octree built from solid z-ranges per column, no N^3 volume) renders the same frame as
    assign_map on the materialised 1024^3 volume."""
    import bench

    scene = bench.bench_scene("c3")
    S = pkg.scene
    a = make_caster(pkg, scene, True, assign_octree=False)
    assert a.compute()
    want, want_aux = a.draw(), a.read_aux()
    a.close()
    cols = S.Scene(scene.n, None, scene.width, scene.height, scene.cam_pos, scene.cam_dir, scene.lights,
                   max_distance=scene.max_distance, columns=S.terrain_columns(scene.n, "shell"))
    b = make_caster(pkg, cols, True)
    assert b.compute(), b.last_error()
    assert np.array_equal(b.draw(), want)
    got_aux = b.read_aux()
    for f in AUX_FIELDS:
        assert np.array_equal(got_aux[f], want_aux[f]), f
    b.close()


def test_deep_tree_4096(pkg, oracle):
    """BASELINE config 4 depth: 4096^3 (12 octree levels = 6 levels of the 64-tree), octree built from columns.
    Checker: the oracle walking the same column table (vr_oracle.h: widened map access), every 40th row."""
    S = pkg.scene
    n = 4096
    lo, hi = S.terrain_columns(n, "shell")
    pos, direction = S.make_camera(n, hi, 9)
    scene = S.Scene(n, None, 1920, 1080, pos, direction, S.make_lights(n), max_distance=3 * n, columns=(lo, hi))
    c = make_caster(pkg, scene, True)
    assert c.stats().levels == 6
    assert c.compute(), c.last_error()
    got, aux = c.draw(), c.read_aux()
    ref_rgba, ref_aux, _ = oracle.raycast(scene, row_stride=40)
    assert_same_frame(ref_rgba[::40], ref_aux[::40], got[::40], aux[::40], "4096^3 rows")
    assert (aux["flags"] & 1).mean() > 0.3 and aux["steps_total"].max() > 2000
    c.close()


def test_view_batch(pkg, oracle):
    """BASELINE config 5 mechanism: a batch of cameras over one scene (vr_compute_views); every frame equals the
    oracle's frame for that camera."""
    import torch

    S = pkg.scene
    n = 64
    vol = S.terrain_map(n, "shell", reflect_fraction=0.03)
    h = S.heightfield(n)
    cams = np.array([np.concatenate([d, p]) for p, d in (S.make_camera(n, h, i) for i in range(6))], dtype=np.float32)
    scene = S.Scene(n, vol, 320, 200, cams[0, 2:].copy(), cams[0, :2].copy(), S.make_lights(n), max_distance=3 * n)
    c = make_caster(pkg, scene, True, assign_octree=False, aux=False)
    out = torch.zeros((len(cams), 200, 320, 4), dtype=torch.uint8, device="cuda:0")
    out[...] = torch.tensor([255, 255, 255, 100], dtype=torch.uint8, device="cuda:0")
    assert c.compute_views(cams, out.data_ptr()) and c.sync()
    frames = out.cpu().numpy()
    for i, cam in enumerate(cams):
        scene.cam_dir[:] = cam[:2]
        scene.cam_pos[:] = cam[2:]
        ref, _, _ = oracle.raycast(scene, want_aux=False)
        assert np.array_equal(frames[i], ref), f"view {i}"
    c.close()


def test_octree_save_load_and_l2_window(pkg, oracle, tmp_path):
    """Octree::Load replacement (ref include/map/Octree.h:38, never defined there): a context that only loads the
    saved octree renders the frame of the context that built it; the L2 access-policy window changes nothing."""
    scene = pkg.scene.make_scene("features-low")
    ref_rgba, _, _ = oracle.raycast(scene, want_aux=False)
    a = make_caster(pkg, scene, True, assign_octree=False, aux=False)
    path = tmp_path / "scene.vr64"
    assert a.octree_save(str(path)) and path.stat().st_size > 64
    assert a.set_option("l2_persist", 1) and a.compute() and np.array_equal(a.draw(), ref_rgba)
    assert a.set_option("l2_persist", 0) and a.compute() and np.array_equal(a.draw(), ref_rgba)
    a.close()
    b = pkg.CUDACaster()
    assert b.init(0)
    assert b.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n)
    assert b.add_to_settings_buffer("using_octree", "OCTENABLED", 0)
    assert b.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance)
    assert b.octree_load(str(path)), b.last_error()
    assert b.assign_camera(scene.cam_dir, scene.cam_pos) and b.create_viewport(scene.width, scene.height)
    assert b.assign_lights(scene.lights) and b.create_texture_atlas(scene.atlas) and b.validate(), b.last_error()
    assert b.set_option("walk", 0) and b.compute() and np.array_equal(b.draw(), ref_rgba)
    bad = tmp_path / "bad.vr64"
    bad.write_bytes(b"VR64" + bytes(40))
    assert not b.octree_load(str(bad)) and "not a valid octree" in b.last_error()
    b.close()


def assert_same_except_ties(ref_rgba, ref_aux, got_rgba, got_aux, what):
    """Per-axis walk (option walk=1): identical to the oracle on every pixel whose ray never takes a multi-axis
    (exact tie) step; on the flagged tie pixels -- rays through a voxel edge, "degenerate" in BASELINE.json's
    sense -- the north_star tolerance applies: same first hit, RGBA max abs diff <= 1."""
    tie = (ref_aux["flags"] & 4) != 0
    ok = ~tie
    for f in AUX_FIELDS:
        a, b = ref_aux[f], got_aux[f]
        if f == "flags":
            a, b = a & 0xFB, b & 0xFB
        bad = np.any(np.atleast_3d(a != b), axis=-1) & ok
        assert not bad.any(), f"{what}: {f} differs on {int(bad.sum())} non-tie pixels, first {np.argwhere(bad)[0]}"
    diff = np.abs(ref_rgba.astype(np.int16) - got_rgba.astype(np.int16)).max(axis=-1)
    assert not (diff[ok] > 0).any(), f"{what}: RGBA differs on non-tie pixels"
    assert tie.mean() < 0.02, f"{what}: {tie.mean():.4f} of the pixels are tie pixels"
    if tie.any():
        # EVERY tie pixel: RGBA within +-1 and the same first hit (voxel, face, type); only the step counts may differ
        assert (diff[tie] <= 1).all(), f"{what}: {int((diff[tie] > 1).sum())} tie pixels beyond +-1"
        for f in ("hit", "face", "hit_type"):
            assert np.array_equal(ref_aux[f][tie], got_aux[f][tie]), f"{what}: first hit ({f}) differs on a tie pixel"
    return int(tie.sum())


@pytest.mark.parametrize("name", SMALL)
def test_axis_walk_small(pkg, oracle, name):
    scene = pkg.scene.make_scene(name)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root))
    c = make_caster(pkg, scene, True)
    assert c.set_option("walk", 1) and c.compute(), c.last_error()
    assert_same_except_ties(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"axis walk {name}")
    c.close()


def test_axis_walk_terrain(pkg, oracle):
    S = pkg.scene
    for n, w, h, cam in ((64, 1280, 720, 3), (256, 1920, 1080, 4)):
        vol = S.terrain_map(n, "shell", reflect_fraction=0.05 if n == 64 else 0.0)
        pos, direction = S.make_camera(n, S.heightfield(n), cam)
        scene = S.Scene(n, vol, w, h, pos, direction, S.make_lights(n), max_distance=3 * n)
        ref_rgba, ref_aux, _ = oracle.raycast(scene)
        c = make_caster(pkg, scene, True, assign_octree=False)
        assert c.set_option("walk", 1) and c.compute(), c.last_error()
        assert_same_except_ties(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"axis walk terrain {n}")
        c.close()


def test_axis_walk_full_size_c3(pkg):
    """1024^3 @3840x2160: per-axis walk vs merged walk (both CUDA): identical except on tie pixels."""
    import bench

    scene = bench.bench_scene("c3")
    c = make_caster(pkg, scene, True, assign_octree=False)
    assert c.compute()
    ref_rgba, ref_aux = c.draw(), c.read_aux()
    assert c.set_option("walk", 1) and c.compute()
    ties = assert_same_except_ties(ref_rgba, ref_aux, c.draw(), c.read_aux(), "axis walk c3")
    assert ties > 0
    c.close()


def _device_tree(pkg, volume, gpu_build):
    """64-tree arrays (nodes as uint32[n, 4], leaf types) of a map, built on the device or on the host."""
    import torch

    c = pkg.CUDACaster()
    assert c.init(0)
    assert c.set_option("gpu_build", 1 if gpu_build else 0)
    assert c.assign_map(volume), c.last_error()
    nb, tb, levels, dim = c.native_tree_info()
    nodes = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
    types = torch.empty(tb, dtype=torch.uint8, device="cuda:0")
    assert c.native_tree_copy(nodes.data_ptr(), types.data_ptr())
    torch.cuda.synchronize()
    st = c.stats()
    out = nodes.cpu().numpy().view(np.uint32).reshape(-1, 4), types.cpu().numpy(), levels, dim, st
    c.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["features", "terrain64", "terrain256-solid", "random32", "empty16", "full8", "dim4", "terrain128-mirrors"])
def test_gpu_builder_equals_host_builder(pkg, kind):
    """vr_build.cu (the on-device replacement of Octree::Generate) emits exactly the arrays of the host builder
    vr_native_from_dense: same nodes in the same breadth-first order, same child pointers, same plane bits, same
    leaf types -- including maps whose root is wider than the map (edge not a power of 4), the empty map (a root
    without children), a full map and transparent voxel values (neither 5 nor 6)."""
    S = pkg.scene
    rng = np.random.default_rng(5)
    if kind == "features":
        vol = S.features_map(32)
    elif kind == "terrain64":
        vol = S.terrain_map(64, "shell")
    elif kind == "terrain256-solid":
        vol = S.terrain_map(256, "solid")
    elif kind == "terrain128-mirrors":
        vol = S.terrain_map(128, "shell", reflect_fraction=0.1)
    elif kind == "random32":
        vol = np.zeros((32, 32, 32), np.int8)
        vol[rng.random(vol.shape) < 0.05] = 5
        vol[rng.random(vol.shape) < 0.02] = 6
        vol[rng.random(vol.shape) < 0.05] = 3
        vol[rng.random(vol.shape) < 0.01] = -7
    elif kind == "empty16":
        vol = np.zeros((16, 16, 16), np.int8)
    elif kind == "full8":
        vol = np.full((8, 8, 8), 5, np.int8)
    else:
        vol = np.zeros((4, 4, 4), np.int8)
        vol[1, 2, 3] = 6
        vol[0, 0, 0] = 5
    g_nodes, g_types, g_levels, g_dim, g_st = _device_tree(pkg, vol, True)
    h_nodes, h_types, h_levels, h_dim, h_st = _device_tree(pkg, vol, False)
    assert (g_levels, g_dim) == (h_levels, h_dim)
    assert g_nodes.shape == h_nodes.shape and np.array_equal(g_nodes, h_nodes)
    assert np.array_equal(g_types, h_types)
    assert g_st.solid_voxels == h_st.solid_voxels == int(np.isin(vol, (5, 6)).sum())
    assert g_st.build_ms > 0 and h_st.build_ms == 0


@pytest.mark.gpu
def test_gpu_builder_full_size_c3(pkg):
    """1024^3 (1 GiB map): the tree built on the device from the dense map, the tree built on the device from the column
    tables and the tree built on the host from the column tables render the same frame; the voxel count is the map's."""
    import bench

    S = pkg.scene
    scene = bench.bench_scene("c3")
    c = make_caster(pkg, scene, True, assign_octree=False, aux=False)       # assign_map -> device build
    st = c.stats()
    assert st.build_ms > 0 and st.solid_voxels == int(np.count_nonzero(scene.volume))
    assert c.set_option("walk", 1) and c.compute()
    got = c.draw().copy()
    c.close()
    lo, hi = S.terrain_columns(scene.n, "shell")
    scene2 = S.Scene(scene.n, None, scene.width, scene.height, scene.cam_pos, scene.cam_dir, scene.lights,
                     max_distance=scene.max_distance, columns=(lo, hi))
    d = make_caster(pkg, scene2, True, assign_octree=False, aux=False)      # assign_columns -> device build from the tables
    assert d.stats().build_ms > 0 and d.stats().native_nodes == st.native_nodes
    assert d.set_option("walk", 1) and d.compute()
    assert np.array_equal(got, d.draw())
    d.close()
    e = pkg.CUDACaster()
    e.load_scene(scene2, use_octree=True, assign_octree=False, walk=1, gpu_build=False)      # ... and on the host
    assert e.stats().build_ms == 0 and e.stats().native_nodes == st.native_nodes
    assert e.compute() and np.array_equal(got, e.draw())
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["features", "features-low", "small"])
@pytest.mark.parametrize("count", [2, 3])
def test_multi_light_extension(pkg, oracle, name, count):
    """LIGHT_COUNT > 1 (extension; BASELINE configs[2] says "2 shadow lights", the reference kernel reads light 0
    only): dense and octree kernels equal the oracle's restatement of the extension on all pixels, and the
    per-axis walk equals it up to the tie tolerance.  LIGHT_COUNT absent or 1 is the reference path."""
    from test_emu_parity import _with_lights

    scene = _with_lights(pkg, pkg.scene.make_scene(name), count)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root), shadow_lights=count)
    one_rgba, one_aux, _ = oracle.raycast(scene, octree=(desc, root))
    for use_octree in (False, True):
        c = pkg.CUDACaster()
        c.load_scene(scene, use_octree=use_octree, shadow_lights=count, walk=0)
        assert c.enable_aux(True) and c.compute(), c.last_error()
        assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"{name} lights={count} octree={use_octree}")
        if use_octree:
            assert c.set_option("walk", 1) and c.compute()
            assert_same_except_ties(ref_rgba, ref_aux, c.draw(), c.read_aux(), f"{name} lights={count} axis walk")
            assert c.set_option("walk", 0)
        # LIGHT_COUNT back to 1: the reference path again (the settings buffer is re-read every frame)
        assert c.overwrite_setting("light_count", 1) and c.compute()
        assert_same_frame(one_rgba, one_aux, c.draw(), c.read_aux(), f"{name} lights back to 1")
        c.close()


@pytest.mark.gpu
def test_multi_light_terrain_256(pkg, oracle):
    S = pkg.scene
    n = 256
    vol = S.terrain_map(n, "shell")
    pos, direction = S.make_camera(n, S.heightfield(n), 10)
    scene = S.Scene(n, vol, 960, 540, pos, direction, S.make_lights(n, 2), max_distance=3 * n)
    ref_rgba, ref_aux, cnt = oracle.raycast(scene, shadow_lights=2, want_counters=True)
    assert cnt["shadow_rays"] == 2 * int(((ref_aux["flags"] & 1) != 0).sum())
    c = pkg.CUDACaster()
    c.load_scene(scene, use_octree=True, assign_octree=False, shadow_lights=2, walk=0)
    assert c.enable_aux(True) and c.compute(), c.last_error()
    assert_same_frame(ref_rgba, ref_aux, c.draw(), c.read_aux(), "terrain 256, 2 lights")
    assert c.set_option("walk", 1) and c.compute()
    assert_same_except_ties(ref_rgba, ref_aux, c.draw(), c.read_aux(), "terrain 256, 2 lights, axis walk")
    c.close()


@pytest.mark.gpu
def test_push_bands_into_registered_host_frame(pkg, oracle):
    """vr_host_register + vr_push_bands with a host target: two band sets (as two ranks would hold them) copied
    device -> host straight into frame order reproduce the single-context frame."""
    import torch

    tiles = pkg.tiles
    scene = pkg.scene.make_scene("features")
    full = make_caster(pkg, scene, True, aux=False)
    assert full.compute()
    want = full.draw().copy()
    full.close()
    H, W = scene.height, scene.width
    lay = tiles.BandLayout(H, W, 8, 2)
    shared = tiles.SharedHostFrame(lay, None, 0, count=1)
    casters = [make_caster(pkg, scene, True, aux=False) for _ in range(2)]
    shared.register(casters[0])
    for rank, c in enumerate(casters):
        assert c.set_bands(8, 2, rank)
        slab = torch.zeros((lay.slab_rows, W, 4), dtype=torch.uint8, device="cuda:0")
        assert c.compute_into(slab.data_ptr()), c.last_error()
        assert c.push_bands(slab.data_ptr(), shared.ptr(0)), c.last_error()
        torch.cuda.synchronize()
    assert np.array_equal(shared.frame(0), want)
    shared.close()                       # unregisters through casters[0]
    for c in casters:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_interleave_renders_in_place(pkg, oracle, world):
    """vr_set_tiles: the tile sets of `world` ranks, rendered one after the other IN PLACE into one full-size frame
    (what the ranks of a multi-GPU run do into the root's frame over NVLink), give the single-context frame --
    dense and octree kernels, frame sizes that are not multiples of the tile or of world."""
    import torch

    scene = pkg.scene.make_scene("features")           # 200 x 120: 6.25 tiles wide
    for use_octree in (False, True):
        full = make_caster(pkg, scene, use_octree, aux=False)
        assert full.compute()
        want = full.draw().copy()
        frame = torch.tensor([255, 255, 255, 100], dtype=torch.uint8, device="cuda:0").repeat(scene.height, scene.width, 1).contiguous()
        covered = torch.zeros((scene.height, scene.width, 4), dtype=torch.uint8, device="cuda:0")
        for rank in range(world):
            assert full.set_tiles(world, rank)
            assert full.compute_into(frame.data_ptr()), full.last_error()
            one = torch.zeros_like(covered)
            assert full.compute_into(one.data_ptr())
            torch.cuda.synchronize()
            assert not ((one != 0) & (covered != 0)).any()          # tile sets are disjoint
            covered |= one
        assert np.array_equal(frame.cpu().numpy(), want)
        assert full.set_bands(8, 2, 1) and not full.compute_into(frame.data_ptr())      # bands + tiles: refused
        assert "mutually exclusive" in full.last_error()
        full.close()


@pytest.mark.gpu
def test_cuda_equals_reference_generated_golden_vectors(pkg):
    """The committed outputs of the reference's own kernel (tests/golden/ref, written by make_golden_ref.py from
    oracle/_ref): RGBA8 and the written mask of the dense and octree CUDA kernels, max_distance 20 (verbatim kernel)
    and the scene's (lifted build)."""
    import pathlib

    from test_oracle import _golden_ref_scene

    g = pathlib.Path(__file__).parent / "golden" / "ref"
    frames = sorted(f for f in g.glob("*.npz") if not f.name.endswith("-octree.npz") and not f.name.startswith("viewport-"))
    assert len(frames) >= 16
    for f in frames:
        z = np.load(f)
        scene = _golden_ref_scene(pkg, str(z["scene"]))
        scene.max_distance = int(z["max_distance"])
        for use_octree in (False, True):
            c = make_caster(pkg, scene, use_octree, aux=False)
            assert c.compute(), c.last_error()
            got = c.draw()
            assert np.array_equal(got, z["rgba"]), f"{f.name} use_octree={use_octree}"
            # unwritten pixels keep the initial fill (255,255,255,100) of create_viewport (host:280-286)
            assert (got[~z["written"]] == np.array([255, 255, 255, 100], np.uint8)).all()
            c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["head", "features", "features-low", "features-high", "features-mirror", "small"])
def test_cuda_equals_reference_kernel_source(pkg, name):
    """Directly against the reference's own kernel source compiled for the CPU (oracle/_ref/libref_kernel_md.so, see
    tests/test_reference_kernel.py), not through the oracle: RGBA8 of the dense kernel and of the octree kernel (merged
    walk) on all pixels, including which pixels are left unwritten."""
    import ref_kernel_lib as R

    if not R.available(True):
        pytest.skip("oracle/_ref/libref_kernel_md.so not built (needs /root/reference at build time)")
    scene = pkg.scene.make_scene(name)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, _ = R.raycast(scene, octree=(desc, root), lifted=True)
    for use_octree in (False, True):
        c = make_caster(pkg, scene, use_octree, aux=False)
        assert c.compute(), c.last_error()
        assert np.array_equal(c.draw(), ref_rgba), f"{name} use_octree={use_octree}"
        c.close()


@pytest.mark.gpu
def test_cuda_random_scenes(pkg, oracle):
    """Differential fuzzing on the GPU (kept last in this file): 24 random scenes (see test_emu_parity.random_scene),
    dense kernel / octree kernel with the merged walk / with the per-axis walk, 1-3 lights, against the oracle."""
    from test_emu_parity import assert_walk_matches, random_scene

    rng = np.random.default_rng(2)
    for it in range(24):
        scene, nl = random_scene(pkg, rng)
        desc, root = pkg.octree_generate(scene.volume)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, octree=(desc, root), shadow_lights=nl)
        for use_octree in (False, True):
            c = pkg.CUDACaster()
            c.load_scene(scene, use_octree=use_octree, shadow_lights=nl, walk=0)
            assert c.enable_aux(True) and c.compute(), c.last_error()
            assert_walk_matches(ref_rgba, ref_aux, c.draw(), c.read_aux(), False, f"scene {it} octree={use_octree}")
            if use_octree:
                assert c.set_option("walk", 1) and c.compute()
                assert_walk_matches(ref_rgba, ref_aux, c.draw(), c.read_aux(), True, f"scene {it} per-axis walk")
            c.close()


@pytest.mark.gpu
def test_octree_load_rejects_corrupt_files(pkg, tmp_path):
    """vr_octree_load (the on-disk format of the never-implemented Octree::Load, include/map/Octree.h:38) validates a
    file before anything of it reaches the device: sizes against the file length before allocating, child pointers
    level by level (inner levels inside the node array and in BFS order, the leaf level inside the type array)."""
    import struct

    scene = pkg.scene.make_scene("features")
    a = make_caster(pkg, scene, True, assign_octree=False, aux=False)
    good = tmp_path / "good.vr64"
    assert a.octree_save(str(good))
    a.close()
    blob = bytearray(good.read_bytes())
    magic, version, dim, levels, nn, nt = struct.unpack_from("<4sIiiQQ", blob, 0)
    assert magic == b"VR64" and nn > 4 and len(blob) == 32 + 16 * nn + nt
    b = pkg.CUDACaster()
    assert b.init(0)

    def rejected(data, what):
        p = tmp_path / "bad.vr64"
        p.write_bytes(bytes(data))
        assert not b.octree_load(str(p)), what
        assert "octree" in b.last_error()

    assert b.octree_load(str(good)), b.last_error()
    rejected(blob[:-7], "truncated")
    rejected(blob + b"\0" * 5, "trailing bytes")
    huge = bytearray(blob)
    struct.pack_into("<Q", huge, 16, (1 << 32) - 1)                       # node count far beyond the file: no 64 GB resize
    rejected(huge, "node count beyond the file")
    ptr = bytearray(blob)
    struct.pack_into("<I", ptr, 32 + 8, nn + 5)                            # root child_base beyond the node array
    rejected(ptr, "inner child pointer out of range")
    leaf = bytearray(blob)
    struct.pack_into("<I", leaf, 32 + 16 * (nn - 1) + 8, nt)               # last leaf: types beyond the type array
    rejected(leaf, "leaf child pointer out of range")
    back = bytearray(blob)
    struct.pack_into("<I", back, 32 + 16 * 1 + 8, 0)                       # a level-1 node pointing back at the root
    rejected(back, "child pointer not in BFS order")
    assert b.octree_load(str(good)), b.last_error()
    b.close()


def _column_tree(pkg, lo, hi, gpu_build, voxel_type=5):
    import torch

    c = pkg.CUDACaster()
    assert c.init(0)
    assert c.set_option("gpu_build", 1 if gpu_build else 0)
    assert c.assign_columns(lo, hi, voxel_type), c.last_error()
    nb, tb, levels, dim = c.native_tree_info()
    nodes = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
    types = torch.empty(tb, dtype=torch.uint8, device="cuda:0")
    assert c.native_tree_copy(nodes.data_ptr(), types.data_ptr())
    torch.cuda.synchronize()
    st = c.stats()
    out = nodes.cpu().numpy().view(np.uint32).reshape(-1, 4), types.cpu().numpy(), levels, dim, float(st.build_ms)
    c.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["shell64", "solid64", "shell256", "shell1024", "solid1024", "irregular32", "empty16", "dim4", "dim8", "shell4096"])
def test_gpu_column_builder_equals_host_builder(pkg, kind):
    """The column-driven device builder (vr_build.cu: vr_build_tree_columns_device -- brick layers per brick column,
    radix sort by hierarchical key, reduce-by-key per level; no N^3 volume, no dense workspace) emits exactly the arrays
    of the host builder vr_native_from_columns, up to the 4096^3 map of BASELINE configs[3]."""
    S = pkg.scene
    rng = np.random.default_rng(21)
    if kind.startswith("shell") or kind.startswith("solid"):
        lo, hi = S.terrain_columns(int(kind[5:]), kind[:5])
    elif kind == "irregular32":
        n = 32
        lo = rng.integers(-3, n, size=(n, n)).astype(np.int32)
        hi = lo + rng.integers(-3, 9, size=(n, n)).astype(np.int32)          # empty columns (hi < lo), some past the top
    elif kind == "empty16":
        lo, hi = np.ones((16, 16), np.int32), np.zeros((16, 16), np.int32)
    else:
        n = int(kind[3:])
        lo = rng.integers(0, n, size=(n, n)).astype(np.int32)
        hi = lo + rng.integers(0, 3, size=(n, n)).astype(np.int32)
    vt = 6 if kind == "irregular32" else 5
    d_nodes, d_types, d_levels, d_dim, ms = _column_tree(pkg, lo, hi, True, vt)
    h_nodes, h_types, h_levels, h_dim, _ = _column_tree(pkg, lo, hi, False, vt)
    assert ms > 0 and (d_levels, d_dim) == (h_levels, h_dim) and d_nodes.shape == h_nodes.shape and d_types.shape == h_types.shape
    assert np.array_equal(d_nodes, h_nodes), f"{kind}: {int((d_nodes != h_nodes).any(axis=1).sum())} nodes differ"
    assert np.array_equal(d_types, h_types)
    print(f"{kind}: {d_nodes.shape[0]} nodes, {d_types.shape[0]} voxel types, device build {ms:.2f} ms")


@pytest.mark.gpu
def test_gl_interop_entry_points_fail_loudly_without_a_gl_context(pkg):
    """draw() with CL/GL sharing (ref src/CLCaster.cpp:330-332, :840-842) -> CUDA-GL interop (vr_gl_register_texture /
    vr_gl_draw).  This image has no OpenGL, so only the failure path can run: registering a texture without a current GL
    context is refused with a message, drawing without a registered texture too, and the context keeps working."""
    scene = pkg.scene.make_scene("head")
    c = make_caster(pkg, scene, False, aux=False)
    assert c.compute()
    before = c.draw().copy()
    assert not c.gl_draw() and "no texture registered" in c.last_error()
    assert not c.gl_register_texture(1) and "gl_register_texture" in c.last_error()
    assert c.gl_unregister()
    assert c.compute() and np.array_equal(c.draw(), before)
    c.close()
