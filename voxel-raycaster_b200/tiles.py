"""Multi-GPU screen-tile scheduler: interleaved row bands, one rank per GPU.

The frame is cut into bands of `band_rows` rows; band b belongs to rank b % world.  Each rank renders its
bands into a compact slab (C ABI: vr_set_bands / vr_compute_into); the slabs are exchanged with one
all_gather (NCCL on GPUs, gloo in the CPU tests) and un-interleaved on rank 0 with one strided copy.
Nothing here renders: it is index arithmetic + the collective, shared by bench.py and the tests.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class BandLayout:
    height: int
    width: int
    band_rows: int
    world: int

    @property
    def nbands(self) -> int:
        return (self.height + self.band_rows - 1) // self.band_rows

    @property
    def max_bands(self) -> int:
        """bands per rank, padded so that every rank's slab has the same shape (all_gather needs that)"""
        return (self.nbands + self.world - 1) // self.world

    @property
    def slab_rows(self) -> int:
        return self.max_bands * self.band_rows

    def bands_of(self, rank: int) -> list[int]:
        return list(range(rank, self.nbands, self.world))

    def rows_of(self, rank: int) -> list[int]:
        """frame rows owned by `rank`, in slab order (matches vr_frame_params band mapping)"""
        return [r for b in self.bands_of(rank) for r in range(b * self.band_rows, min(self.height, (b + 1) * self.band_rows))]

    def local_rows(self, rank: int) -> int:
        return len(self.rows_of(rank))


def gather_frame(layout: BandLayout, slab, gathered, frame, dist=None, rank: int = 0):
    """slab [slab_rows, W, 4] of this rank -> frame [max_bands*world*band_rows, W, 4] on every rank that passes
    `frame` (rows >= height are padding).  `gathered` is a scratch tensor [world * slab_rows, W, 4] (rank-major concatenation of the slabs)."""
    if layout.world == 1:
        frame.copy_(slab)
        return frame
    dist.all_gather_into_tensor(gathered, slab)
    if frame is not None:
        mb, w, br, W = layout.max_bands, layout.world, layout.band_rows, layout.width
        # band b lives at gathered[b % world, b // world]: one strided copy restores frame order
        frame.view(mb, w, br, W, 4).copy_(gathered.view(w, mb, br, W, 4).permute(1, 0, 2, 3, 4))
    return frame
