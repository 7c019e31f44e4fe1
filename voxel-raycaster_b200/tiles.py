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


class SharedHostFrame:
    """`count` frames [max_bands*world*band_rows, W, 4] uint8 in one POSIX shared-memory file mapped by every rank
    of the node.  With the mapping page-locked (CUDACaster.host_register) each rank copies its own bands device ->
    host into frame order (CUDACaster.push_bands with the mapping as target): the frame reaches host memory over all
    PCIe links at once and no rank ever holds the whole frame on its GPU."""

    def __init__(self, layout: BandLayout, dist=None, rank: int = 0, count: int = 2, directory: str = "/dev/shm"):
        import mmap
        import os

        import numpy as np

        self.layout, self.rank, self.count = layout, rank, count
        self.rows = layout.max_bands * layout.world * layout.band_rows
        self.frame_bytes = self.rows * layout.width * 4
        names = [None]
        if rank == 0:
            names[0] = os.path.join(directory, f"vr_frame_{os.getpid()}_{id(self) & 0xffff:x}")
            with open(names[0], "wb") as f:
                f.truncate(self.frame_bytes * count)
        if dist is not None and layout.world > 1:
            dist.broadcast_object_list(names, src=0)
        self.path = names[0]
        self._fd = os.open(self.path, os.O_RDWR)
        self._map = mmap.mmap(self._fd, self.frame_bytes * count)
        self.array = np.frombuffer(self._map, dtype=np.uint8).reshape(count, self.rows, layout.width, 4)
        self.registered_by = None
        if dist is not None and layout.world > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)              # every rank holds its mapping; the name is no longer needed

    def ptr(self, i: int) -> int:
        return self.array[i].ctypes.data

    def register(self, caster) -> None:
        if not caster.host_register(self.array.ctypes.data, self.frame_bytes * self.count):
            raise RuntimeError(caster.last_error())
        self.registered_by = caster

    def frame(self, i: int):
        """frame i in frame order, height rows"""
        return self.array[i, : self.layout.height]

    def close(self) -> None:
        import os

        if self.registered_by is not None:
            self.registered_by.host_unregister(self.array.ctypes.data)
            self.registered_by = None
        self.array = None
        try:
            self._map.close()
        except BufferError:
            pass
        os.close(self._fd)


class _RawCuda:
    """uint8 device memory owned elsewhere, exposed to torch via the CUDA array interface"""

    def __init__(self, ptr: int, shape) -> None:
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class FramePipeline:
    """Double-buffered multi-GPU frame loop: the gather (+ un-interleave, + optional device->host copy on rank 0)
    of frame k runs on a communication stream while the compute stream already renders frame k+1 into the
    other slab.  `render(ptr)` must enqueue the rendering of this rank's bands into device memory at `ptr`
    on the CURRENT torch stream (CUDACaster.compute_into after set_stream)."""

    def __init__(self, layout: BandLayout, device, dist, rank: int, render, host_frame=None, caster=None):
        """caster != None selects the copy-engine gather: every rank pushes its slab into the root's frame buffer
        through a CUDA-IPC mapping (CUDACaster.push_bands), followed by a one-element all_reduce as the "frame
        complete" signal; otherwise the slabs go through NCCL all_gather + one un-interleave copy."""
        import torch

        self.torch, self.layout, self.dist, self.rank, self.render, self.host_frame = torch, layout, dist, rank, render, host_frame
        self.caster = caster
        self.shared_host = None       # set_shared_host(): every rank copies its bands straight to the host frame
        W = layout.width
        self.slabs = [torch.zeros((layout.slab_rows, W, 4), dtype=torch.uint8, device=device) for _ in range(2)]
        self.gathered = torch.empty((layout.world * layout.slab_rows, W, 4), dtype=torch.uint8, device=device)
        self.frame = torch.empty((layout.max_bands * layout.world * layout.band_rows, W, 4), dtype=torch.uint8, device=device) if rank == 0 else None
        self.comm = torch.cuda.Stream(device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.frames = None
        if caster is not None:
            # two frame buffers on the root, mapped into every other rank
            if rank == 0:
                # cudaMalloc'ed by the caster (torch's caching allocator memory is not reliably IPC-exportable),
                # wrapped as torch tensors through __cuda_array_interface__
                shape = tuple(self.frame.shape)
                nbytes = self.frame.numel()
                self.frame_ptrs = [caster.device_alloc(nbytes) for _ in range(2)]
                self.frames = [torch.as_tensor(_RawCuda(p, shape), device=device) for p in self.frame_ptrs]
                self.frame = self.frames[0]
                handles = [caster.ipc_get_handle(p) for p in self.frame_ptrs]
            else:
                handles = [None, None]
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                self.frame_ptrs = [caster.ipc_open_handle(h) for h in handles]
        self.rendered = [torch.cuda.Event() for _ in range(2)]
        self.collected = [torch.cuda.Event() for _ in range(2)]
        self.k = 0

    def step(self) -> None:
        torch = self.torch
        i = self.k & 1
        cur = torch.cuda.current_stream()
        if self.k >= 2:
            cur.wait_event(self.collected[i])            # slab i has left for the gather of frame k-2
        self.render(self.slabs[i].data_ptr())
        self.rendered[i].record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.rendered[i])
            if self.shared_host is not None:
                # end-to-end path with a HOST result: this rank's bands -> frame order in shared pinned memory
                if not self.shared_caster.push_bands(self.slabs[i].data_ptr(), self.shared_host.ptr(i % self.shared_host.count), self.comm.cuda_stream):
                    raise RuntimeError(self.shared_caster.last_error())
            elif self.caster is not None:
                if not self.caster.push_bands(self.slabs[i].data_ptr(), self.frame_ptrs[i], self.comm.cuda_stream):
                    raise RuntimeError(self.caster.last_error())
                self.dist.all_reduce(self.token)          # completes on the root only after every push has landed
                if self.rank == 0:
                    self.frame = self.frames[i]
            else:
                gather_frame(self.layout, self.slabs[i], self.gathered, self.frame, self.dist, self.rank)
            if self.host_frame is not None and self.rank == 0:
                self.host_frame.copy_(self.frame[: self.layout.height], non_blocking=True)
            self.collected[i].record(self.comm)
        self.k += 1

    def set_shared_host(self, shared_host, caster) -> None:
        self.shared_host, self.shared_caster = shared_host, caster

    def drain(self) -> None:
        """make the current stream wait for every gather issued so far"""
        self.torch.cuda.current_stream().wait_stream(self.comm)
