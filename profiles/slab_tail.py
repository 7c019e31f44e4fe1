"""Tail effect of the 1/8-frame slab kernels (what each rank of an 8-GPU run executes), measured on ONE GPU:
times the 8 band sets of the C3 frame one after the other and compares their sum with the full-frame kernel.

    python profiles/slab_tail.py            # prints per-slab kernel ms, their sum, the full-frame ms
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

pkg = bench.package()
S = pkg.scene
n = 1024
lo, hi = S.terrain_columns(n, "shell")
pos, direction = S.make_camera(n, S.heightfield(n), bench.BENCH_CAMERA)
scene = S.Scene(n, None, 3840, 2160, pos, direction, S.make_lights(n, 1), max_distance=3 * n, columns=(lo, hi))
c = pkg.CUDACaster()
c.load_scene(scene, use_octree=True, assign_octree=False)
assert c.set_option("walk", 1)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lay = pkg.tiles.BandLayout(scene.height, scene.width, bench.BAND_ROWS, world)
slab = torch.zeros((lay.slab_rows, scene.width, 4), dtype=torch.uint8, device="cuda:0")


stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
assert c.set_stream(stream.cuda_stream)


def timed(reps=20):
    for _ in range(3):
        assert c.compute_into(slab.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        assert c.compute_into(slab.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


per = []
for r in range(world):
    assert c.set_bands(bench.BAND_ROWS, world, r)
    per.append(timed())
assert c.set_bands(bench.BAND_ROWS, 1, 0)
full_slab = torch.zeros((lay.max_bands * world * bench.BAND_ROWS, scene.width, 4), dtype=torch.uint8, device="cuda:0")
slab = full_slab
full = timed()
tiles = []
for r in range(world):
    assert c.set_tiles(world, r)
    tiles.append(timed())
assert c.set_tiles(1, 0)
print(f"2-D tile interleave {[round(v, 4) for v in tiles]} max {max(tiles):.4f} sum {sum(tiles):.4f} -> scaling at {world} GPUs = {full / max(tiles):.2f}x")
print(f"slabs {[round(v, 4) for v in per]} max {max(per):.4f} sum {sum(per):.4f} full {full:.4f} -> scaling at {world} GPUs = {full / max(per):.2f}x")
