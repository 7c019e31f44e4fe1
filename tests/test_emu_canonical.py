"""CPU tests of the closed-form walk (voxel-raycaster_b200/csrc/vr_canon.h, library default walk = 2) compiled for the host
(tests/host_emu): over the undirected top grid (use_svo = 3) and over the directed ones (use_svo = 4, one table per
direction octant of the ray: vr_octree.cpp: vr_native_grid_directed) against Oracle-B = the oracle with closed-form
crossing times.  Same bar as the GPU suite (tests/test_gpu_canonical.py: assert_equals_oracle_b): identical on every
pixel whose ray makes no exact multi-axis step, same first hit and RGBA8 within +-1 on the others.  Plus the invariants of
the directed grids against the dense map."""
import numpy as np
import pytest

import emu_lib
from conftest import oracle_bias
from test_gpu_canonical import assert_equals_oracle_b
from test_gpu_parity import SMALL


@pytest.mark.parametrize("name", SMALL)
def test_closed_form_walk_small_scenes(pkg, oracle, name):
    scene = pkg.scene.make_scene(name)
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root), canonical_t=True)
    bias = oracle_bias(oracle, scene, desc, root)
    frames = []
    for use_svo in (3, 4):
        rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo)
        assert_equals_oracle_b(ref_rgba, ref_aux, rgba, aux, f"{name} svo={use_svo}")
        frames.append((rgba, aux))
    # the two grids only change which empty cells a ray is handed: the frames are the same, the lookups fewer
    assert np.array_equal(frames[0][0], frames[1][0])
    assert frames[1][1]["lookups"].sum() <= frames[0][1]["lookups"].sum()


def test_closed_form_walk_terrain(pkg, oracle):
    """64^3 shell terrain with 5 % mirrors, three cameras, two lights: the directed grids hand out far larger cells (rays
    leaving the surface towards a light), the frame must not change."""
    S = pkg.scene
    vol = S.terrain_map(64, "shell", reflect_fraction=0.05)
    fewer = 0
    for cam in (1, 3, 5):
        pos, direction = S.make_camera(64, S.heightfield(64), cam)
        scene = S.Scene(64, vol, 320, 180, pos, direction, S.make_lights(64, 2), max_distance=192)
        table = oracle.make_ray_table(scene.width, scene.height)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, table, shadow_lights=2, canonical_t=True)
        look = []
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, use_svo=use_svo, shadow_lights=2)
            assert_equals_oracle_b(ref_rgba, ref_aux, rgba, aux, f"terrain cam {cam} svo={use_svo}")
            look.append(int(aux["lookups"].sum()))
        fewer += look[1] < look[0]
    assert fewer == 3


def test_closed_form_walk_random_and_sparse_scenes(pkg, oracle):
    """Differential fuzzing on the CPU: random scenes (maps 8^3..64^3, cameras inside / outside / on integer coordinates,
    1-3 lights) and sparse 128^3 maps with cameras in collapsed empty cells (negative start bias, kernel:353)."""
    from test_emu_parity import random_scene
    from test_gpu_canonical import sparse_scene

    rng = np.random.default_rng(21)
    cases = [random_scene(pkg, rng) for _ in range(24)] + [sparse_scene(pkg, rng, 128) for _ in range(4)]
    for it, (scene, nl) in enumerate(cases):
        table = oracle.make_ray_table(scene.width, scene.height)
        desc, root = pkg.octree_generate(scene.volume)
        ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
        bias = oracle_bias(oracle, scene, desc, root)
        if scene.n < 8:
            continue                                   # no top grid below 8^3: the library falls back to walk 0
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_equals_oracle_b(ref_rgba, ref_aux, rgba, aux, f"scene {it} (n={scene.n}) svo={use_svo}")


@pytest.mark.parametrize("kind", ["terrain64", "terrain256", "random32", "sparse128", "half32"])
def test_directed_grid_invariants(pkg, kind):
    """vr_native_grid_directed: in every octant table a block is non-empty exactly when it holds a set voxel (same node
    entry as the undirected grid); an empty entry describes a box that starts at the block, extends along the octant's
    direction of travel, lies inside the map and holds no set voxel; a cube entry is maximal (one more layer of blocks
    would leave the map, hit a voxel or exceed the cap) and at least as wide as the undirected grid's centred cube."""
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
    und, g, bits = emu_lib.grid_from_tree(nodes, levels, n)
    grid, g2, bits2 = emu_lib.grid_from_tree(nodes, levels, n, directed=True)
    G = 1 << bits
    assert (g2, bits2) == (g, bits) and grid.shape == (8, G, G, G)
    bs = 1 << g
    cap = min(64, G)
    solid = (vol == 5) | (vol == 6)
    blocks = solid.reshape(G, bs, G, bs, G, bs).any(axis=(1, 3, 5))                      # [bz, by, bx]
    occ = np.zeros((G + 1, G + 1, G + 1), np.int64)
    empties = np.argwhere(~blocks)
    if len(empties) > 6000:
        empties = empties[rng.choice(len(empties), 6000, replace=False)]
    wider = 0
    for o in range(8):
        t = grid[o]
        assert np.array_equal(t[blocks], und[blocks]) and ((t & 0x80000000) != 0).sum() == blocks.sum()
        # mirrored occupancy + summed-area table: axis a flipped when bit a of the octant is set (x = bit 0 = last axis)
        mb = blocks
        if o & 1: mb = mb[:, :, ::-1]
        if o & 2: mb = mb[:, ::-1, :]
        if o & 4: mb = mb[::-1, :, :]
        occ[1:, 1:, 1:] = mb.astype(np.int64).cumsum(0).cumsum(1).cumsum(2)

        def count(lo, hi):          # non-empty blocks in [lo, hi) per axis, mirrored block coordinates (z, y, x)
            (z0, y0, x0), (z1, y1, x1) = lo, hi
            return (occ[z1, y1, x1] - occ[z0, y1, x1] - occ[z1, y0, x1] - occ[z1, y1, x0]
                    + occ[z0, y0, x1] + occ[z0, y1, x0] + occ[z1, y0, x0] - occ[z0, y0, x0])

        for bz, by, bx in empties:
            e = int(t[bz, by, bx])
            m, ext = (1 << (e & 31)) - 1, e >> 8
            k = (G - 1 - bz if o & 4 else bz, G - 1 - by if o & 2 else by, G - 1 - bx if o & 1 else bx)   # mirrored (z, y, x)
            assert m >= bs - 1 and ext % bs == 0 and (ext == 0 or m == bs - 1), (kind, o, e)
            lo = [((kk << g) & ~m) >> g for kk in k]
            hi = [((((kk << g) | m) + ext) >> g) + 1 for kk in k]
            assert min(lo) >= 0 and max(hi) <= G, (kind, o, bz, by, bx, e)
            assert count(lo, hi) == 0, (kind, o, bz, by, bx, e)
            if m == bs - 1:                                                             # a cube with the block in its rear corner
                edge = ext // bs + 1
                assert lo == list(k) and edge <= cap
                if edge < cap:                                                          # maximal
                    grown = [kk + edge + 1 for kk in k]
                    assert max(grown) > G or count(k, grown) > 0, (kind, o, bz, by, bx, e)
                u = int(und[bz, by, bx])
                if (u & 31) == g:                                                       # undirected: centred cube of radius r
                    assert edge >= min((u >> 8) // bs + 1, cap)
                    wider += edge > (u >> 8) // bs + 1
    if kind not in ("random32", "half32"):             # (half32: the aligned 16^3 cells reach further everywhere)
        assert wider > 0


def test_division_by_constants_is_the_ieee_quotient():
    """vr_div_const (vr_trace.h): texel / 255 and steps / 700 through the reciprocal + two FMAs give the correctly rounded
    quotient of kernel:654-656 / 565 / 716 for every operand they can see (0..255, 0..2^24): exhaustive."""
    import ctypes as C

    fn = emu_lib.lib().emu_div_const_mismatches
    fn.restype = C.c_long
    assert fn() == 0


def _solid_volumes(pkg):
    S = pkg.scene
    rng = np.random.default_rng(12)
    out = {"solid64": S.terrain_map(64, "solid"), "solid128-mirrors": S.terrain_map(128, "solid", reflect_fraction=0.002)}
    v = np.zeros((32, 32, 32), np.int8)
    v[:17] = 5
    v[17:21, 8:24, 8:24] = 6                       # a slab of the other type on top: solid nodes of both types
    out["half32-two-types"] = v
    out["full16"] = np.full((16, 16, 16), 5, np.int8)      # the root itself collapses
    v = S.terrain_map(64, "solid").copy()
    holes = rng.random(v.shape) < 0.01
    v[holes & (v != 0)] = 0                        # 1 % holes: only some bricks stay solid
    out["solid64-holes"] = v
    return out


@pytest.mark.parametrize("kind", ["solid64", "solid128-mirrors", "half32-two-types", "full16", "solid64-holes"])
def test_solid_collapse_structure(pkg, kind):
    """vr_native_collapse_solid (VR_NODE_SOLID): the collapsed tree answers every point query like the uncollapsed one,
    is smaller, keeps BFS order (children of a node contiguous, levels in order) and holds no node below a solid node."""
    vol = _solid_volumes(pkg)[kind]
    n = vol.shape[0]
    emu_lib.set_collapse(False)
    nodes0, types0, levels0 = emu_lib.tree_from_dense(vol)
    emu_lib.set_collapse(True)
    nodes, types, levels = emu_lib.tree_from_dense(vol)
    assert levels == levels0 and len(nodes) <= len(nodes0) and len(types) < len(types0)
    SOLID = 0x80000000
    mask = nodes[:, 0].astype(np.uint64) | (nodes[:, 1].astype(np.uint64) << np.uint64(32))
    base = nodes[:, 2]
    solid = (base & SOLID) != 0
    assert solid.any() and (mask[solid] == np.uint64(0xFFFFFFFFFFFFFFFF)).all() and np.isin(base[solid] & 0xFF, (5, 6)).all()
    # walk the levels: children of non-solid inner nodes are contiguous and consecutive across the level (BFS order)
    lo, hi = 0, 1
    for l in range(levels):
        leaf = l == levels - 1
        nxt = hi
        for i in range(lo, hi):
            if solid[i]:
                continue
            pc = bin(int(mask[i])).count("1")
            if leaf:
                assert int(base[i]) + pc <= len(types)
            elif pc:
                assert int(base[i]) == nxt
                nxt += pc
        lo, hi = hi, nxt
        if lo == hi:
            break
    assert hi == len(nodes) or lo == hi

    def query(nd, ty, x, y, z):
        idx, s = 0, 2 * (levels - 1)
        while True:
            m = int(nd[idx, 0]) | (int(nd[idx, 1]) << 32)
            b = int(nd[idx, 2])
            if b & SOLID:
                return b & 0xFF
            ci = ((x >> s) & 3) | (((y >> s) & 3) << 2) | (((z >> s) & 3) << 4)
            if not (m >> ci) & 1:
                return 0
            rank = bin(m & ((1 << ci) - 1)).count("1")
            if s == 0:
                return int(ty[b + rank])
            idx, s = b + rank, s - 2

    rng = np.random.default_rng(3)
    for x, y, z in rng.integers(0, n, size=(3000, 3)):
        want = int(vol[z, y, x])
        x, y, z = int(x), int(y), int(z)
        assert query(nodes, types, x, y, z) == want == query(nodes0, types0, x, y, z), (kind, x, y, z)


@pytest.mark.parametrize("kind", ["solid64", "solid128-mirrors", "half32-two-types", "solid64-holes"])
def test_all_walks_over_collapsed_trees(pkg, oracle, kind):
    """Every octree walk of the device core (host build) over a tree with collapsed solid subtrees: walk 0 == the oracle's
    dense DDA on every pixel, walk 1 except ties, the closed-form walk == Oracle-B over both kinds of top grid -- and the
    frames equal those over the uncollapsed tree."""
    from test_emu_parity import assert_walk_matches

    S = pkg.scene
    vol = _solid_volumes(pkg)[kind]
    n = vol.shape[0]
    if kind.startswith("solid"):
        cams = [S.make_camera(n, S.heightfield(n), 3)]
    else:
        h = np.full((n, n), 16, np.int32)
        h[8:24, 8:24] = 20
        cams = [S.make_camera(n, h, 1), S.make_camera(n, h, 2)]      # one looks at the mirror slab, one at the ground
    seen = 0.0
    for pos, direction in cams:
        scene = S.Scene(n, vol, 256, 144, pos, direction, S.make_lights(n, 2), max_distance=3 * n)
        table = oracle.make_ray_table(scene.width, scene.height)
        a_rgba, a_aux, _ = oracle.raycast(scene, table, shadow_lights=2)
        b_rgba, b_aux, _ = oracle.raycast(scene, table, shadow_lights=2, canonical_t=True)
        frames = {}
        for collapse in (True, False):
            emu_lib.set_collapse(collapse)
            for use_svo in (1, 2, 3, 4):
                rgba, aux = emu_lib.raycast(scene, table, use_svo=use_svo, shadow_lights=2)
                if use_svo <= 2:
                    assert_walk_matches(a_rgba, a_aux, rgba, aux, use_svo == 2, f"{kind} collapse={collapse} svo={use_svo}")
                else:
                    assert_equals_oracle_b(b_rgba, b_aux, rgba, aux, f"{kind} collapse={collapse} svo={use_svo}")
                frames[(collapse, use_svo)] = rgba
        emu_lib.set_collapse(True)
        for use_svo in (1, 2, 3, 4):
            assert np.array_equal(frames[(True, use_svo)], frames[(False, use_svo)]), use_svo
        seen = max(seen, float(((a_aux["flags"] & 3) != 0).mean()))
    assert seen > 0.3                                     # not mostly sky: lit or reflected hits


def _fuzz_case(pkg, seed, it, kind):
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "fuzz"))
    import fuzz_closed_form as F

    F.KINDS = [kind] if kind else "random,random,sparse,terrain,tunnel".split(",")      # (the campaign's default mix)
    return F, F.make_case(seed, it)


def test_fuzz_finding_skipped_next_light_reports_the_hit_iteration(pkg, oracle):
    """tests/fuzz/fuzz_closed_form.py, scene (202, 902): a pixel whose SECOND light lies exactly in a coordinate plane of the
    hit point is skipped (kernel:671 applied to the multi-light extension).  The device core reports the step counter of the
    hit's iteration for it, as both do for a first light; the oracle used to report one more.  All walks == oracle on every
    pixel, including that counter."""
    from test_emu_parity import assert_walk_matches

    F, (kind, scene, nl, collapse) = _fuzz_case(pkg, 202, 902, "tunnel")
    assert nl == 2
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    ref_rgba, ref_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl)
    b_rgba, b_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
    skipped = (ref_aux["status"] == oracle.ST_SKIP_REDIRECT) & ((ref_aux["flags"] & oracle.FL_LIT) != 0)
    assert skipped.any(), "the scene holds a pixel skipped at a light"
    bias = oracle_bias(oracle, scene, desc, root)
    emu_lib.set_collapse(collapse)
    try:
        for use_svo in (0, 1, 2):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_walk_matches(ref_rgba, ref_aux, rgba, aux, use_svo == 2, f"svo={use_svo}")
            assert np.array_equal(aux["steps_total"][skipped], ref_aux["steps_total"][skipped])
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_equals_oracle_b(b_rgba, b_aux, rgba, aux, f"svo={use_svo}")
    finally:
        emu_lib.set_collapse(True)


def test_fuzz_finding_unobserved_ties_inside_a_cell(pkg, oracle):
    """The closed-form walk sees a multi-axis step only where it looks: at the step that leaves an empty cell and inside
    bricks.  An exact tie strictly inside a cell is walked as two steps, so distance_traveled is one higher than Oracle-B's
    (which steps voxel by voxel).  Visible on tie pixels only: fog (RGBA +-1), and -- scene (302, 250) -- a shadow ray that ends
    by max_distance (kernel:357) one step short of the voxel Oracle-B's still reaches (alpha of a shadowed vs a lit pixel);
    scene (301, 103): the two kinds of top grid hand out different cells, so their frames differ by 1 on such a pixel.
    First hit and face are the same in all of them; on that pixel of (302, 250) the walk agrees with the REFERENCE walk (oracle A)."""
    F, (kind, scene, nl, collapse) = _fuzz_case(pkg, 302, 250, "sparse")
    table = oracle.make_ray_table(scene.width, scene.height)
    desc, root = pkg.octree_generate(scene.volume)
    a_rgba, a_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl)
    b_rgba, b_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
    bias = oracle_bias(oracle, scene, desc, root)
    emu_lib.set_collapse(collapse)
    try:
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            problems, ties, far = F.classify(b_rgba, b_aux, rgba, aux)
            assert not problems, problems
            assert far == 1
            y, x = np.argwhere(np.abs(rgba.astype(int) - b_rgba.astype(int)).max(-1) > 1)[0]
            assert (b_aux["flags"][y, x] & 4) and aux["steps_total"][y, x] == b_aux["steps_total"][y, x] + 1
            assert np.array_equal(rgba[y, x], a_rgba[y, x]) and aux["steps_total"][y, x] == a_aux["steps_total"][y, x]
        F2, (kind, scene, nl, collapse) = _fuzz_case(pkg, 301, 103, "sparse")
        table = oracle.make_ray_table(scene.width, scene.height)
        desc, root = pkg.octree_generate(scene.volume)
        b_rgba, b_aux, _ = oracle.raycast(scene, table, octree=(desc, root), shadow_lights=nl, canonical_t=True)
        bias = oracle_bias(oracle, scene, desc, root)
        emu_lib.set_collapse(collapse)
        frames = []
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(scene, table, bias=bias, use_svo=use_svo, shadow_lights=nl)
            assert_equals_oracle_b(b_rgba, b_aux, rgba, aux, f"svo={use_svo}")
            frames.append(rgba)
        d = np.abs(frames[0].astype(int) - frames[1].astype(int)).max(-1)
        tie = (b_aux["flags"] & 4) != 0
        assert (d > 0).sum() == 1 and d.max() == 1 and not (d[~tie] > 0).any()
    finally:
        emu_lib.set_collapse(True)


def test_fuzz_finding_camera_on_a_voxel_corner(pkg, oracle):
    """tests/fuzz/fuzz_closed_form.py, scene (1002, 4830): camera at integer coordinates, max_distance 5.  intersection_t
    starts at 0 on all three axes, so the first step of every primary ray moves along all of them at once and counts ONCE
    (kernel:558, 714).  Inside the camera's empty cell the closed-form walk would not see that tie, count three steps and
    end every ray (max_distance, kernel:357) before the voxel the reference still reaches: the first cell of a primary ray
    is therefore the start voxel alone when the camera sits on a voxel edge (vr_canon.h: on_edge).  Every pixel and every
    counter equals Oracle-B's -- and the same for a camera on an edge (two integer coordinates) of a terrain map."""
    F, (kind, scene, nl, collapse) = _fuzz_case(pkg, 1002, 4830, None)
    assert np.array_equal(scene.cam_pos, np.floor(scene.cam_pos)) and scene.max_distance == 5
    S = pkg.scene
    n = 64
    pos, direction = S.make_camera(n, S.heightfield(n), 2)
    pos = np.array([np.floor(pos[0]), np.floor(pos[1]), pos[2]], np.float32)
    edge = S.Scene(n, S.terrain_map(n, "shell"), 160, 96, pos, direction, S.make_lights(n, 1), max_distance=20)
    # (5002, 2093): a corner camera inside a collapsed empty octree cell.  The start bias (kernel:353) differs between the
    # axes, so the tie of the integer axes is not the first step: such frames are traced voxel by voxel (vr_cam_on_edge = 2)
    _, (_, biased, nl_b, _) = _fuzz_case(pkg, 5002, 2093, None)
    assert np.array_equal(biased.cam_pos, np.floor(biased.cam_pos))
    emu_lib.set_collapse(collapse)
    try:
        for sc, lights in ((scene, nl), (edge, 1), (biased, nl_b)):
            table = oracle.make_ray_table(sc.width, sc.height)
            desc, root = pkg.octree_generate(sc.volume)
            b_rgba, b_aux, _ = oracle.raycast(sc, table, octree=(desc, root), shadow_lights=lights, canonical_t=True)
            assert ((b_aux["flags"] & 4) != 0).mean() > 0.9, "(nearly) every ray starts with a multi-axis step"
            bias = oracle_bias(oracle, sc, desc, root)
            assert (sc is biased) == any(b != 0 for b in bias)
            for use_svo in (3, 4):
                rgba, aux = emu_lib.raycast(sc, table, bias=bias, use_svo=use_svo, shadow_lights=lights)
                assert np.array_equal(rgba, b_rgba), f"svo={use_svo}"
                for f in ("hit", "face", "status", "hit_type", "steps_first", "steps_total"):
                    assert np.array_equal(aux[f], b_aux[f]), (use_svo, f)
    finally:
        emu_lib.set_collapse(True)
