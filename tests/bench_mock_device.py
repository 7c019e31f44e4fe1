"""TEST INFRASTRUCTURE (CPU only): runs bench.py's control flow -- argument handling, the timed regions, the collectives of the
N > 1 path, the JSON line -- against a FAKE device, so that a mistake in that plumbing (an undefined name, a wrong key, a
collective one rank skips) is found here and not on the GPU box at the end of a round.  Nothing is rendered and no number it
prints means anything: the caster is a stub that counts calls, CUDA events read a virtual clock that every fake launch
advances by half a millisecond, `torch.device("cuda", i)` is the CPU and the process group is gloo.

    python tests/bench_mock_device.py [bench.py flags]            (RANK / WORLD_SIZE / MASTER_* from the environment)

Used by tests/test_bench_host.py; never imported by the product or by bench.py.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

CLOCK = {"ms": 0.0}
LAUNCH_MS = 0.5


class FakeEvent:
    def __init__(self, enable_timing: bool = False) -> None:
        self.t = None

    def record(self, stream=None) -> None:
        self.t = CLOCK["ms"]

    def elapsed_time(self, other: "FakeEvent") -> float:
        assert self.t is not None and other.t is not None, "elapsed_time of an event that was never recorded"
        return max(other.t - self.t, 1e-6)


class FakeStream:
    cuda_stream = 0

    def __init__(self, device=None) -> None:
        pass


fake_cuda = types.SimpleNamespace(
    is_available=lambda: True, set_device=lambda i: None, synchronize=lambda *a: None, Stream=FakeStream, set_stream=lambda s: None,
    Event=FakeEvent, current_stream=lambda *a: FakeStream())


class TorchProxy(types.ModuleType):
    """`import torch` inside bench.main() resolves to this: torch with cuda replaced and device("cuda", i) mapped to the CPU"""

    def __init__(self) -> None:
        super().__init__("torch")

    def __getattr__(self, name):
        if name == "cuda":
            return fake_cuda
        if name == "device":
            return lambda *a, **k: torch.device("cpu")
        return getattr(torch, name)


_real_init = dist.init_process_group


def _init_gloo(backend=None, **kw):
    kw.pop("device_id", None)
    return _real_init("gloo", **kw)


dist.init_process_group = _init_gloo
torch.Tensor.pin_memory = lambda self, *a, **k: self

real_pkg = importlib.import_module("voxel-raycaster_b200")


class FakeCaster:
    """counts calls; every method the bench does not look at returns True"""

    MGPU_HOST_FRAME = 1

    def __init__(self) -> None:
        self.launches = 0
        self.issued = 0
        self.w = self.h = 0
        self.calls: dict[str, int] = {}

    def __getattr__(self, name):
        def anything(*a, **k):
            self.calls[name] = self.calls.get(name, 0) + 1
            return True
        return anything

    def last_error(self) -> str:
        return "fake"

    def create_viewport(self, w, h, *a) -> bool:
        self.w, self.h = int(w), int(h)
        return True

    def _launch(self) -> None:
        self.launches += 1
        CLOCK["ms"] += LAUNCH_MS

    def compute_into(self, ptr) -> bool:
        self._launch()
        return True

    def frame_begin(self) -> bool:
        self._launch()
        return True

    def frame_end(self):
        return np.zeros((self.h, self.w, 4), dtype=np.uint8)

    def stats(self):
        return types.SimpleNamespace(kernel_launches=self.launches, native_nodes=1, native_bytes=16, levels=3, build_ms=1.0,
                                     build_masks_ms=0.5, last_kernel_ms=LAUNCH_MS, bias=(0, 0, 0))

    def read_aux(self):
        a = np.zeros((self.h, self.w), dtype=[("status", "u1"), ("flags", "u1"), ("node_fetches", "<u4"), ("lookups", "<u4"), ("steps_total", "<u4")])
        a["status"] = 1
        a["flags"][::2] = 1
        return a

    def native_tree_info(self):
        return 16, 1, 3, 64

    # the multi-GPU scheduler
    def mgpu_frame(self) -> int:
        self._launch()
        self.issued += 1
        return self.issued - 1

    def mgpu_frame_wait(self, k: int):
        assert 0 <= k < self.issued, "waited for a frame that was not issued"
        return self.host.ctypes.data                   # (device frames go through FakeRaw, which ignores the address)

    def mgpu_init(self, session, world, rank, flags=0) -> bool:
        self.issued = 0
        self.host = np.zeros((self.h, self.w, 4), dtype=np.uint8)
        return True


class FakeRaw(np.ndarray):
    def __new__(cls, ptr, shape):
        return np.zeros(shape, dtype=np.uint8).view(cls)


fake_tiles = types.SimpleNamespace(**{k: getattr(real_pkg.tiles, k) for k in dir(real_pkg.tiles) if not k.startswith("__")})
fake_tiles._RawCuda = FakeRaw
fake_pkg = types.SimpleNamespace(CUDACaster=FakeCaster, tiles=fake_tiles, scene=real_pkg.scene, octree_generate=real_pkg.octree_generate)
bench.package = lambda: fake_pkg

if __name__ == "__main__":
    sys.modules["torch"] = TorchProxy()
    sys.argv = [str(ROOT / "bench.py")] + sys.argv[1:]
    bench.main()
