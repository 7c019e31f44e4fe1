"""Generates tests/golden/ref/*.npz from the REFERENCE'S OWN CODE executed on the CPU (oracle/_ref/*.so, built by
`make -C oracle ref` from the sources under /root/reference; see tests/test_reference_kernel.py).  Unlike the files
next to it (regression fixtures written by the oracle), these are reference outputs: committed so that the pin also
holds where oracle/_ref cannot be rebuilt.

    python tests/golden/make_golden_ref.py
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import ref_kernel_lib as R  # noqa: E402

pkg = importlib.import_module("voxel-raycaster_b200")
S = pkg.scene
out = HERE / "ref"
out.mkdir(exist_ok=True)
assert R.available(False) and R.available(True) and R.octree_available(), "run `make -C oracle ref` first"


def terrain64():
    n = 64
    vol = S.terrain_map(n, "shell", reflect_fraction=0.05)
    pos, direction = S.make_camera(n, S.heightfield(n), 3)
    return S.Scene(n, vol, 256, 144, pos, direction, S.make_lights(n), max_distance=3 * n, name="terrain64-cam3")


scenes = [(S.make_scene(n), n) for n in ("head", "tiny", "small", "features", "features-low", "features-high", "features-mirror")]
scenes.append((terrain64(), "terrain64-cam3"))
for scene, name in scenes:
    ref = R.RefOctree(scene.volume)                      # the reference's own Octree::Generate
    for lifted in (False, True):
        rgba, written = R.raycast(scene, octree=(ref.descriptors, ref.root_index), lifted=lifted)
        np.savez_compressed(out / f"{name}-{'lifted' if lifted else 'verbatim'}.npz", scene=name, lifted=lifted,
                            max_distance=scene.max_distance if lifted else 20, rgba=rgba, written=written)
    if name in ("head", "features", "tiny"):
        used = np.flatnonzero(ref.descriptors)
        np.savez_compressed(out / f"{name}-octree.npz", scene=name, root_index=ref.root_index, first_used=int(used.min()),
                            descriptors=ref.descriptors[used.min():])
    ref.close()
    print(name, "ok")

# the ray table as written by the loop of CLCaster::create_viewport (src/CLCaster.cpp:244-275, compiled from the reference's
# source by `make -C oracle ref`): a small even-sized and an odd-sized table in full, the 4K table as a strided sample
if R.viewport_available():
    t4k = R.create_viewport_table(3840, 2160)
    np.savez_compressed(out / "viewport-tables.npz", t64x36=R.create_viewport_table(64, 36), t5x7=R.create_viewport_table(5, 7),
                        t3840x2160_every_60th_row_40th_col=t4k[::60, ::40].copy())
    print("viewport tables ok")
