"""Generates tests/golden/*.npz from the CPU oracle (self-generated regression fixtures: the reference has
no golden vectors of its own and cannot run here -- see oracle/vr_oracle.h "PARITY UNPINNED").

    python tests/golden/make_golden.py
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import oracle_lib as O  # noqa: E402

pkg = importlib.import_module("voxel-raycaster_b200")

for name in ("head", "features", "features-low", "features-high", "features-mirror"):
    scene = pkg.scene.make_scene(name)
    desc, root = pkg.octree_generate(scene.volume)
    rgba, aux, cnt = O.raycast(scene, octree=(desc, root), want_counters=True)
    np.savez_compressed(HERE / f"{name}.npz", scene=name, rgba=rgba, hit=aux["hit"], face=aux["face"], status=aux["status"],
                        flags=aux["flags"], hit_type=aux["hit_type"], steps_first=aux["steps_first"], steps_total=aux["steps_total"])
    print(name, rgba.shape, cnt)

# multi-light extension (LIGHT_COUNT > 1; defined by the oracle's next_light)
from test_emu_parity import _with_lights  # noqa: E402

for name, count in (("features", 2), ("features-low", 3)):
    scene = _with_lights(pkg, pkg.scene.make_scene(name), count)
    desc, root = pkg.octree_generate(scene.volume)
    rgba, aux, cnt = O.raycast(scene, octree=(desc, root), want_counters=True, shadow_lights=count)
    np.savez_compressed(HERE / f"{name}-lights{count}.npz", scene=name, lights=count, rgba=rgba, hit=aux["hit"], face=aux["face"],
                        status=aux["status"], flags=aux["flags"], hit_type=aux["hit_type"], steps_first=aux["steps_first"],
                        steps_total=aux["steps_total"])
    print(name, count, rgba.shape, cnt)
