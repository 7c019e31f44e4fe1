"""CPU tests of the multi-GPU tile scheduler: world_size 2 and 3 over gloo.  Each rank fills its band slab with
the ORACLE's pixels for the rows it owns (standing in for vr_compute_into on a GPU), the slabs go through the
same all_gather + un-interleave code bench.py uses, and rank 0 must end up with the full oracle frame."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, band_rows, frame_np, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = importlib.import_module("voxel-raycaster_b200").tiles
    H, W = frame_np.shape[:2]
    lay = tiles.BandLayout(H, W, band_rows, world)
    slab = torch.zeros((lay.slab_rows, W, 4), dtype=torch.uint8)
    rows = lay.rows_of(rank)
    slab[: len(rows)] = torch.from_numpy(frame_np[rows])
    gathered = torch.empty((world * lay.slab_rows, W, 4), dtype=torch.uint8)
    frame = torch.empty((lay.max_bands * world * band_rows, W, 4), dtype=torch.uint8) if rank == 0 else None
    tiles.gather_frame(lay, slab, gathered, frame, dist, rank)
    if rank == 0:
        np.save(out_path, frame[:H].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,band_rows", [(2, 8), (3, 7), (2, 16)])
def test_band_gather_reassembles_oracle_frame(pkg, oracle, tmp_path, world, band_rows):
    scene = pkg.scene.make_scene("features-low")
    ref, _, _ = oracle.raycast(scene, want_aux=False)
    out = tmp_path / "frame.npy"
    port = 29600 + world * 10 + band_rows
    mp.spawn(_worker, args=(world, port, band_rows, ref, str(out)), nprocs=world, join=True)
    assert np.array_equal(np.load(out), ref)


def test_band_layout_matches_c_abi_mapping(pkg):
    """BandLayout.rows_of must agree with the kernel's local-row -> frame-row map (vr_types.h)."""
    tiles = pkg.tiles
    for H, br, world in ((2160, 8, 8), (120, 7, 2), (50, 16, 3), (9, 4, 4)):
        lay = tiles.BandLayout(H, 1, br, world)
        seen = []
        for rank in range(world):
            rows = lay.rows_of(rank)
            # kernel formula: ly -> ((ly // br) * world + rank) * br + ly % br, dropped when >= H
            want = [((ly // br) * world + rank) * br + ly % br for ly in range(lay.slab_rows)]
            assert rows == [y for y in want if y < H]
            seen += rows
        assert sorted(seen) == list(range(H))


def _shm_worker(rank, world, port, band_rows, frame_np, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = importlib.import_module("voxel-raycaster_b200").tiles
    H, W = frame_np.shape[:2]
    lay = tiles.BandLayout(H, W, band_rows, world)
    shared = tiles.SharedHostFrame(lay, dist, rank, count=2)
    assert shared.array.shape == (2, lay.max_bands * world * band_rows, W, 4)
    # what vr_push_bands does with the mapping as target: band j of the slab -> frame band j * world + rank
    rows = lay.rows_of(rank)
    slab = np.zeros((lay.slab_rows, W, 4), np.uint8)
    slab[: len(rows)] = frame_np[rows]
    view = shared.array[1].reshape(lay.max_bands, world, band_rows, W, 4)
    view[:, rank] = slab.reshape(lay.max_bands, band_rows, W, 4)
    dist.barrier()
    if rank == 0:
        np.save(out_path, shared.frame(1).copy())
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,band_rows", [(2, 8), (3, 5)])
def test_shared_host_frame_assembles_in_frame_order(pkg, oracle, tmp_path, world, band_rows):
    """The N > 1 end-to-end path: every rank writes its bands into one shared-memory host frame (on the GPU box with
    a device->host 2-D copy, here with numpy using the same band -> frame-row map); rank 0 reads the whole frame."""
    scene = pkg.scene.make_scene("features-low")
    ref, _, _ = oracle.raycast(scene, want_aux=False)
    out = tmp_path / "frame.npy"
    mp.spawn(_shm_worker, args=(world, 29700 + world * 10 + band_rows, band_rows, ref, str(out)), nprocs=world, join=True)
    assert np.array_equal(np.load(out), ref)
