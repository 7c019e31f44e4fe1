"""PINS THE ORACLE TO THE REFERENCE'S OWN CODE (CPU tests).

The reference has no tests or golden vectors and no OpenCL runtime exists in this image, but its sources can still
be EXECUTED: `make -C oracle ref` compiles, from where they lie under /root/reference,
  * kernels/ray_caster_kernel.cl  -- through oracle/ref_shim/cl_shim.h (OpenCL C types / operators / built-ins in
    C++; the only edit, made on the fly, is the vector-literal syntax `(typeN)(` -> `typeN(`), once verbatim
    (max_distance 20, kernel:326) and once with that one constant read from a variable;
  * src/map/Octree.cpp + include/util.hpp -- with stand-ins for the three SFML headers they include,
into oracle/_ref/*.so (git-ignored; they travel to the GPU box, the reference sources do not).  These tests require
the oracle's restatement to reproduce that code bit for bit: every RGBA8 pixel and the written/skipped mask of the
kernel, every entry of Octree::Generate's 100 000-descriptor buffer and its root index, util.hpp's Normalize.
What stays defined by this repo rather than by the reference: the OpenCL built-ins the kernel calls (normalize,
fast_length, read_imagef, ...), which any OpenCL runtime would supply and which cl_shim.h defines exactly like the
oracle does (IEEE binary32, oracle/vr_oracle.h)."""
import numpy as np
import pytest

import ref_kernel_lib as R

needs_kernel = pytest.mark.skipif(not (R.available(False) and R.available(True)),
                                  reason="oracle/_ref/libref_kernel*.so not built (needs /root/reference: make -C oracle ref)")
needs_octree = pytest.mark.skipif(not R.octree_available(), reason="oracle/_ref/libref_octree.so not built")

SCENES = ["head", "tiny", "small", "features", "features-low", "features-high", "features-mirror"]


def _same(oracle_rgba, oracle_aux, ref_rgba, ref_written, what):
    diff = np.abs(oracle_rgba.astype(np.int16) - ref_rgba.astype(np.int16)).max(-1)
    assert not diff.any(), f"{what}: {int((diff > 0).sum())} pixels differ from the reference kernel (max {int(diff.max())})"
    skipped = (oracle_aux["status"] == 0) | (oracle_aux["status"] == 4)         # kernel:293 / :671 / :694 `return`s
    assert np.array_equal(ref_written, ~skipped), f"{what}: written-pixel mask differs"


@needs_kernel
@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("lifted", [False, True])
def test_oracle_equals_reference_kernel(pkg, oracle, name, lifted):
    """Dense branch (OCTENABLED != 0) of the reference kernel, all pixels: HEAD's scene, shadow hits, lit pixels,
    reflections with the +1 voxel_step accident (kernel:698), out-of-map rays, a camera in a collapsed empty octree
    cell (non-zero get_oct_vox bias, kernel:353) -- verbatim with max_distance 20 and with the scene's max_distance."""
    scene = pkg.scene.make_scene(name)
    desc, root = pkg.octree_generate(scene.volume)
    md = scene.max_distance if lifted else 20
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root), max_distance=md)
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted=lifted)
    _same(o_rgba, o_aux, r_rgba, written, f"{name} lifted={lifted}")
    if lifted and name in ("features", "features-low"):
        assert (o_aux["status"] == 3).any() and ((o_aux["flags"] & 1) != 0).any()    # shadowed and lit pixels were compared


@needs_kernel
@pytest.mark.parametrize("cam", [1, 3, 4])
def test_oracle_equals_reference_kernel_terrain(pkg, oracle, cam):
    """64^3 terrain with 5 % mirror voxels at 640x360, three cameras, max_distance 3N."""
    S = pkg.scene
    n = 64
    vol = S.terrain_map(n, "shell", reflect_fraction=0.05)
    pos, direction = S.make_camera(n, S.heightfield(n), cam)
    scene = S.Scene(n, vol, 640, 360, pos, direction, S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root))
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted=True)
    _same(o_rgba, o_aux, r_rgba, written, f"terrain64 cam {cam}")
    assert ((o_aux["flags"] & 2) != 0).any()


@needs_kernel
def test_oracle_equals_reference_kernel_256(pkg, oracle):
    """256^3 (the deepest octree the kernel's 8-entry stacks allow, kernel:119-124), 480x270, max_distance 768."""
    S = pkg.scene
    n = 256
    vol = S.terrain_map(n, "shell")
    pos, direction = S.make_camera(n, S.heightfield(n), 10)
    scene = S.Scene(n, vol, 480, 270, pos, direction, S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root))
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted=True)
    _same(o_rgba, o_aux, r_rgba, written, "terrain256")


@needs_kernel
def test_oracle_equals_reference_kernel_random_values(pkg, oracle):
    """Voxel values other than 0/5/6 are transparent (kernel:575); sparse random map, camera in open space."""
    S = pkg.scene
    rng = np.random.default_rng(3)
    n = 16
    vol = np.zeros((n, n, n), np.int8)
    vol[rng.random((n, n, n)) < 0.03] = 5
    vol[rng.random((n, n, n)) < 0.01] = 6
    vol[rng.random((n, n, n)) < 0.02] = 3
    vol[8, 8, 8] = 0
    scene = S.Scene(n, vol, 96, 64, np.array([8.4, 8.6, 8.3], np.float32), np.array([1.9, 0.7], np.float32),
                    S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root))
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted=True)
    _same(o_rgba, o_aux, r_rgba, written, "random16")


@needs_octree
@pytest.mark.parametrize("name", ["head", "tiny", "small", "features"])
def test_octree_generate_equals_reference(pkg, oracle, name):
    """Octree::Generate (src/map/Octree.cpp:13-43, 171-323), executed: all 100 000 buffer entries and the root index
    equal the oracle's restated generator; the reference's own Validate accepts it; and the kernel renders the same
    frame from the reference's buffer as from the product's differently laid out buffer."""
    scene = pkg.scene.make_scene(name)
    ref = R.RefOctree(scene.volume)
    mine, my_root, used = oracle.octree_generate(scene.volume, 100000)
    assert ref.root_index == my_root
    assert np.array_equal(ref.descriptors, mine)
    assert ref.validate()
    if name == "head":            # the known answer the oracle tests derive from the sources
        assert (used, ref.root_index, int(ref.descriptors[ref.root_index])) == (585, 99415, 0x00FF0001)
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(0, scene.n, size=(300, 3)):
        found, _, _ = ref.get_voxel(int(x), int(y), int(z))
        assert bool(found) == bool(scene.volume[z, y, x])
        assert bool(oracle.get_oct_vox(mine, my_root, scene.n, (x, y, z))[0]) == bool(found)
    if R.available(True):
        a, _ = R.raycast(scene, octree=(ref.descriptors, ref.root_index), lifted=True)
        desc, root = pkg.octree_generate(scene.volume)
        b, _ = R.raycast(scene, octree=(desc, root), lifted=True)
        assert np.array_equal(a, b)
    ref.close()


@needs_octree
def test_ray_table_uses_reference_normalize(oracle):
    """create_viewport (src/CLCaster.cpp:244-275) cannot be compiled here (OpenCL / GL), but the one non-trivial
    function it calls can: util.hpp's Normalize, applied to the double-rotated ray.  The oracle's table equals it."""
    w, h = 64, 36
    table = oracle.make_ray_table(w, h)
    s157, c157 = np.sin(1.57), np.cos(1.57)
    for y in range(-h // 2, h // 2, 5):
        for x in range(-w // 2, w // 2, 7):
            rx, ry, rz = np.float32(-800.0), np.float32(x), np.float32(y)
            v = np.array([np.float32(float(rz) * s157 + float(rx) * c157), ry, np.float32(float(rz) * c157 - float(rx) * s157)], np.float32)
            want = R.normalize(v)
            got = table[y + h // 2, x + w // 2, :3] if table.ndim == 3 else table.reshape(h, w, 4)[y + h // 2, x + w // 2, :3]
            assert np.array_equal(want.view(np.uint32), np.asarray(got, np.float32).view(np.uint32)), (x, y)


@pytest.mark.skipif(not R.viewport_available(), reason="oracle/_ref/libref_viewport.so not built")
@pytest.mark.parametrize("size", [(64, 36), (50, 50), (1280, 720), (3840, 2160), (5, 7)])
def test_ray_table_equals_reference_create_viewport_loop(oracle, size):
    """The loop of CLCaster::create_viewport itself (src/CLCaster.cpp:244-275), cut out of the reference's source and compiled
    here (oracle/ref_shim/ref_viewport_host.cpp): the oracle's table equals it bit for bit on every ray, including the
    4K table of the headline workload.  Odd sizes: the reference's loops run over 2 * (n / 2) rows / columns, the last row
    / column stays as `new sf::Vector4f[]` left it (zero); the oracle restates that too."""
    w, h = size
    want = R.create_viewport_table(w, h)
    got = oracle.make_ray_table(w, h)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert np.isfinite(got).all() and (got[: 2 * (h // 2), : 2 * (w // 2), :3] != 0).any(axis=-1).all()


@needs_kernel
@needs_octree
@pytest.mark.parametrize("name", ["head", "features", "tiny"])
def test_reference_octree_branch_never_registers_a_hit(pkg, name):
    """Why the reference's octree branch (OCTENABLED == 0, kernel:359-550, "not working") is REPLACED rather than
    ported: executed as written, on the reference's own octree buffer, it never assigns voxel_data (kernel:549 is
    commented out), so no pixel is textured or lit -- every pixel is an opaque flat grey (1 - steps/8, kernel:372-377)
    whatever the scene.  The caster instead renders, with OCTENABLED == 0, what the dense branch renders."""
    scene = pkg.scene.make_scene(name)
    ref = R.RefOctree(scene.volume)
    rgba, written = R.raycast(scene, octree=(ref.descriptors, ref.root_index), lifted=False, octenabled=0)
    ref.close()
    assert written.all()
    assert (rgba[..., 3] == 255).all()
    assert (rgba[..., 0] == rgba[..., 1]).all() and (rgba[..., 1] == rgba[..., 2]).all()
    assert len(np.unique(rgba[..., 0])) <= 8
    dense, _ = R.raycast(scene, octree=(ref.descriptors, ref.root_index), lifted=False, octenabled=1)
    assert not np.array_equal(dense, rgba)


@needs_kernel
@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_equals_reference_kernel_random_scenes(pkg, oracle, seed):
    """Differential fuzzing: 40 random scenes per seed -- 8^3..32^3 maps of random density with mirrors and transparent
    values, cameras at random / integer / half-integer coordinates and outside the map, random and axis-aligned view
    directions (zero ray components: skipped pixels), random lights, max_distance 5 / 20 / 3N -- all pixels and the
    written mask identical.  (1150 such scenes were run when this check was introduced: no difference.)"""
    S = pkg.scene
    rng = np.random.default_rng(seed)
    for it in range(40):
        n = int(rng.choice([8, 16, 32]))
        vol = np.zeros((n, n, n), np.int8)
        dens = rng.choice([0.01, 0.05, 0.2, 0.6])
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
        lights = np.zeros((8, 10), np.float32)
        lights[0] = [rng.random(), rng.random(), rng.random(), rng.random() * 2, *(rng.random(3) * n * 1.2 - 0.1 * n), -1, -1, -1.5]
        if rng.random() < 0.1:
            lights[0, 4:7] = np.floor(lights[0, 4:7])
        scene = S.Scene(n, vol, 48, 32, pos, d, lights, max_distance=int(rng.choice([20, 3 * n, 5])))
        desc, root = pkg.octree_generate(vol)
        o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root))
        r_rgba, written = R.raycast(scene, octree=(desc, root), lifted=True)
        _same(o_rgba, o_aux, r_rgba, written, f"seed {seed} scene {it} (n {n}, mode {mode})")


needs_wide = pytest.mark.skipif(not R.available("wide"), reason="oracle/_ref/libref_kernel_wide.so not built")


@needs_wide
def test_oracle_equals_widened_reference_kernel_512(pkg, oracle):
    """Beyond 256^3 the reference kernel's 8-entry private stacks overflow (kernel:119-124); the `wide` build widens
    exactly those three arrays to 32 entries (sed, oracle/Makefile) and lifts max_distance.  512^3 terrain (octree
    depth 9), 480x270, max_distance 1536: all pixels identical."""
    S = pkg.scene
    n = 512
    vol = S.terrain_map(n, "shell")
    pos, direction = S.make_camera(n, S.heightfield(n), 9)
    scene = S.Scene(n, vol, 480, 270, pos, direction, S.make_lights(n), max_distance=3 * n)
    desc, root = pkg.octree_generate(vol)
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root))
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted="wide")
    _same(o_rgba, o_aux, r_rgba, written, "terrain512 wide")
    assert o_aux["steps_total"].max() > 300
    # the widened build is the same kernel where the stacks suffice
    small = S.make_scene("features")
    d2, r2 = pkg.octree_generate(small.volume)
    a, wa = R.raycast(small, octree=(d2, r2), lifted="wide")
    b, wb = R.raycast(small, octree=(d2, r2), lifted=True)
    assert np.array_equal(a, b) and np.array_equal(wa, wb)


@needs_wide
def test_oracle_equals_widened_reference_kernel_full_size_c3(pkg, oracle):
    """BASELINE's headline workload itself (1024^3 shell terrain, 3840x2160, max_distance 3072, the bench camera):
    every 32nd row = 261 120 pixels, rays of up to 2 300 steps, oracle == the reference's own (widened) kernel."""
    import bench

    scene = bench.bench_scene("c3")
    desc, root = pkg.octree_generate(scene.volume)
    stride = 32
    o_rgba, o_aux, _ = oracle.raycast(scene, octree=(desc, root), row_stride=stride)
    r_rgba, written = R.raycast(scene, octree=(desc, root), lifted="wide", row_stride=stride)
    rows = slice(0, scene.height, stride)
    _same(o_rgba[rows], o_aux[rows], r_rgba[rows], written[rows], "c3 sampled rows")
    assert o_aux["steps_total"][rows].max() > 2000 and ((o_aux["flags"][rows] & 1) != 0).mean() > 0.5


@needs_octree
def test_octree_generators_on_random_volumes(pkg, oracle):
    """Random 8^3..32^3 volumes of random density: the oracle's restated generator reproduces Octree::Generate's whole
    buffer; the product's generator (different layout) and the reference's answer every point query alike."""
    rng = np.random.default_rng(4)
    for it in range(12):
        n = int(rng.choice([8, 16, 32]))
        vol = (rng.random((n, n, n)) < rng.choice([0.002, 0.05, 0.3, 0.9])).astype(np.int8) * int(rng.choice([5, 6, 3]))
        ref = R.RefOctree(vol)
        buf, root, used = oracle.octree_generate(vol, 100000)
        assert root == ref.root_index and np.array_equal(buf, ref.descriptors), it
        mine, my_root = pkg.octree_generate(vol)
        for x, y, z in rng.integers(0, n, size=(200, 3)):
            a = oracle.get_oct_vox(buf, root, n, (x, y, z))
            b = oracle.get_oct_vox(mine, my_root, n, (x, y, z))
            assert a[:3] == b[:3], (it, x, y, z)
            assert bool(ref.get_voxel(int(x), int(y), int(z))[0]) == bool(a[0]) == bool(vol[z, y, x])
        ref.close()
