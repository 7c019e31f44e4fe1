"""GPU test of the closed-form walk with the camera on a voxel edge or corner (vr_canon.h: vr_canon_first_step_tie,
vr_types.h: vr_cam_on_edge), added in the third session of round 2 when no GPU time was left: its CPU twin
(tests/test_emu_canonical.py::test_fuzz_finding_camera_on_a_voxel_corner) runs the same device code compiled for the host.
The file sorts last so that, under `pytest -x`, it runs after every test that has been seen green on a B200."""
import numpy as np
import pytest

from test_gpu_canonical import INT_FIELDS

pytestmark = pytest.mark.gpu


def test_camera_on_a_voxel_corner_or_edge(pkg, oracle):
    """A camera with integer coordinates: every primary ray starts with a multi-axis step, which counts once (kernel:558,
    714; vr_canon.h: vr_canon_first_step_tie).  Scene (1002, 4830) of tests/fuzz/fuzz_closed_form.py (corner, max_distance
    5: counted per axis, every ray would end before the voxel the reference reaches) and a terrain camera on an edge:
    RGBA8 and every integer field equal Oracle-B's on all pixels, over both kinds of top grid.  CPU twin:
    tests/test_emu_canonical.py::test_fuzz_finding_camera_on_a_voxel_corner."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent / "fuzz"))
    import fuzz_closed_form as F

    F.KINDS = "random,random,sparse,terrain,tunnel".split(",")
    kind, corner, nl, collapse = F.make_case(1002, 4830)
    assert np.array_equal(corner.cam_pos, np.floor(corner.cam_pos)) and corner.max_distance == 5
    S = pkg.scene
    n = 64
    pos, direction = S.make_camera(n, S.heightfield(n), 2)
    pos = np.array([np.floor(pos[0]), np.floor(pos[1]), pos[2]], np.float32)
    edge = S.Scene(n, S.terrain_map(n, "shell"), 160, 96, pos, direction, S.make_lights(n, 1), max_distance=20)
    # (5002, 2093): a corner camera inside a collapsed empty octree cell (start bias, kernel:353): traced voxel by voxel
    _, biased, nl_b, _ = F.make_case(5002, 2093)
    assert np.array_equal(biased.cam_pos, np.floor(biased.cam_pos))
    for scene, lights in ((corner, nl), (edge, 1), (biased, nl_b)):
        desc, root = pkg.octree_generate(scene.volume)
        b_rgba, b_aux, _ = oracle.raycast(scene, octree=(desc, root), shadow_lights=lights, canonical_t=True)
        c = pkg.CUDACaster()
        c.load_scene(scene, use_octree=True, shadow_lights=lights)
        assert c.enable_aux(True)
        for directed in (1, 0):
            assert c.set_option("directed_grid", directed) and c.compute(), c.last_error()
            rgba, aux = c.draw(), c.read_aux()
            assert np.array_equal(rgba, b_rgba), f"directed={directed}"
            for f in INT_FIELDS:
                assert np.array_equal(aux[f], b_aux[f]), (directed, f)
        c.close()
