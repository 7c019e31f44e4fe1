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

    def __init__(self, layout: BandLayout, device, dist, rank: int, render, host_frame=None, caster=None, direct=False):
        """caster != None selects the copy-engine gather: every rank pushes its slab into the root's frame buffer
        through a CUDA-IPC mapping (CUDACaster.push_bands), followed by a one-element all_reduce as the "frame
        complete" signal; otherwise the slabs go through NCCL all_gather + one un-interleave copy.
        direct=True (needs caster, and the caster in tile mode: CUDACaster.set_tiles): there are no slabs at all --
        `render(ptr)` is handed the root's frame itself (local on the root, IPC-mapped elsewhere) and the render
        kernels store their pixels in place over NVLink; three frame buffers, so that a rank may start frame k+3
        as soon as frame k+1 is complete everywhere (no per-frame bubble)."""
        import torch

        self.torch, self.layout, self.dist, self.rank, self.render, self.host_frame = torch, layout, dist, rank, render, host_frame
        self.caster = caster
        self.direct = bool(direct and caster is not None)
        self.shared_host = None       # set_shared_host(): every rank copies its bands straight to the host frame
        self.nbuf = 3 if self.direct else 2
        W = layout.width
        self.slabs = [] if self.direct else [torch.zeros((layout.slab_rows, W, 4), dtype=torch.uint8, device=device) for _ in range(2)]
        self.gathered = torch.empty((layout.world * layout.slab_rows, W, 4), dtype=torch.uint8, device=device) if caster is None else None
        self.frame = torch.empty((layout.max_bands * layout.world * layout.band_rows, W, 4), dtype=torch.uint8, device=device) if (rank == 0 and caster is None) else None
        self.comm = torch.cuda.Stream(device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.frames = None
        if caster is not None:
            # frame buffers on the root, mapped into every other rank
            shape = (layout.max_bands * layout.world * layout.band_rows, W, 4)
            nbytes = shape[0] * shape[1] * shape[2]
            if rank == 0:
                # cudaMalloc'ed by the caster (torch's caching allocator memory is not reliably IPC-exportable),
                # wrapped as torch tensors through __cuda_array_interface__
                self.frame_ptrs = [caster.device_alloc(nbytes) for _ in range(self.nbuf)]
                self.frames = [torch.as_tensor(_RawCuda(p, shape), device=device) for p in self.frame_ptrs]
                self.frame = self.frames[0]
                handles = [caster.ipc_get_handle(p) for p in self.frame_ptrs]
            else:
                handles = [None] * self.nbuf
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                self.frame_ptrs = [caster.ipc_open_handle(h) for h in handles]
        self.rendered = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.collected = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.k = 0
        # alternate_streams(): consecutive frames are rendered on two streams, so that the first CTAs of frame k+1
        # fill the SMs the tail of frame k leaves idle (a 1/8-frame kernel is only ~7 waves of CTAs)
        self.render_streams = None
        self.use_stream = None

    def alternate_streams(self, streams, use_stream) -> None:
        """streams: two torch streams; use_stream(s): make the renderer launch on s (CUDACaster.set_stream)"""
        self.render_streams, self.use_stream = streams, use_stream

    def step(self) -> None:
        torch = self.torch
        i = self.k % self.nbuf
        cur = torch.cuda.current_stream()
        if self.render_streams is not None:
            cur = self.render_streams[self.k & 1]
            self.use_stream(cur)
        if self.k >= 2:
            # buffer i was last used by step k - nbuf; every rank (and the root's consumer) is past it once the
            # collection of step k - 2 has completed here
            cur.wait_event(self.collected[(self.k - 2) % self.nbuf])
        if self.direct:
            self.render(self.frame_ptrs[i])               # pixels go straight into the root's frame (NVLink stores)
            self.rendered[i].record(cur)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.rendered[i])
                self.dist.all_reduce(self.token)          # completes on the root only after every kernel has finished
                if self.rank == 0:
                    self.frame = self.frames[i]
                    if self.host_frame is not None:
                        self.host_frame.copy_(self.frame[: self.layout.height], non_blocking=True)
                self.collected[i].record(self.comm)
            self.k += 1
            return
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
        if self.render_streams is not None:
            for s in self.render_streams:
                self.torch.cuda.current_stream().wait_stream(s)
