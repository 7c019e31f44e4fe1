"""GPU tests of the multi-GPU frame scheduler behind the C ABI (voxel-raycaster_b200/csrc/vr_mgpu.cu: vr_mgpu_*).

world = 1 runs in the standard -m gpu suite (one GPU): the scheduler path (tile mode in place / band mode + strided
copy into shared pinned host memory, device-stored completion counters, frame ring with release) must reproduce
vr_compute's frame bit for bit.  world = 2 needs two GPUs (gpurun --gpus 2): two processes, the octree broadcast with
NCCL from rank 0, frames assembled on the root GPU over NVLink and in shared host memory."""
import ctypes as C
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _device_frame(ptr: int, h: int, w: int) -> np.ndarray:
    import torch

    class Raw:
        __cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

    return torch.as_tensor(Raw(), device="cuda:0").cpu().numpy()


@pytest.mark.parametrize("host_frame", [False, True])
def test_scheduler_world_1(pkg, host_frame):
    scene = pkg.scene.make_scene("features")
    ref = pkg.CUDACaster()
    ref.load_scene(scene, use_octree=True)
    assert ref.compute()
    want = ref.draw()
    ref.close()
    c = pkg.CUDACaster()
    c.load_scene(scene, use_octree=True)
    flags = c.MGPU_HOST_FRAME if host_frame else 0
    assert c.mgpu_init(f"t1_{os.getpid()}_{int(host_frame)}", 1, 0, flags), c.last_error()
    assert c.mgpu_broadcast_octree(), c.last_error()
    assert c.mgpu_barrier(), c.last_error()              # CPU rendezvous of the ranks (a world of one: returns at once)
    h, w = scene.height, scene.width
    for i in range(7):                                   # more frames than the ring holds
        k = c.mgpu_frame()
        assert k == i, c.last_error()
        ptr = c.mgpu_frame_wait(k)
        assert ptr, c.last_error()
        got = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h, w, 4)).copy() if host_frame else _device_frame(ptr, h, w)
        assert np.array_equal(got, want), f"frame {k}"
        assert c.mgpu_frame_release(k)
    assert c.mgpu_shutdown()
    assert c.compute() and np.array_equal(c.draw(), want)          # the context is a plain caster again
    c.close()


WORKER = textwrap.dedent("""
    import ctypes as C, importlib, os, sys
    import numpy as np
    import torch                      # first: the scheduler then uses PyTorch's NCCL, not a second copy
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    rank, world, session, host = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    import bench
    pkg = importlib.import_module("voxel-raycaster_b200")
    S = pkg.scene
    n = 256
    scene = S.Scene(n, S.terrain_map(n, "shell") if rank == 0 else None, 640, 360, *S.make_camera(n, S.heightfield(n), 4), S.make_lights(n), max_distance=3 * n)
    c = pkg.CUDACaster()
    assert c.init(rank)
    assert c.add_to_settings_buffer("octree_dimensions", "OCTDIM", n) and c.add_to_settings_buffer("using_octree", "OCTENABLED", 0)
    assert c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", 3 * n)
    if rank == 0:
        assert c.assign_map(scene.volume)
    assert c.assign_camera(scene.cam_dir, scene.cam_pos) and c.create_viewport(640, 360, 56.25, 90.0)
    assert c.assign_lights(scene.lights) and c.create_texture_atlas(scene.atlas, (16, 16))
    assert c.mgpu_init(session, world, rank, host), c.last_error()
    assert c.mgpu_broadcast_octree(), c.last_error()
    assert c.validate(), c.last_error()
    assert c.mgpu_barrier(), c.last_error()
    frames = []
    for i in range(5):
        k = c.mgpu_frame(); assert k == i, c.last_error()
        ptr = c.mgpu_frame_wait(k); assert ptr is not None, c.last_error()
        if rank == 0:
            if host:
                frames.append(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(360, 640, 4)).copy())
            else:
                class Raw:
                    __cuda_array_interface__ = {{"shape": (360, 640, 4), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}}
                frames.append(torch.as_tensor(Raw(), device="cuda:0").cpu().numpy())
            assert c.mgpu_frame_release(k)
    assert c.mgpu_shutdown()
    if rank == 0:
        # the single-GPU frame of the same scene on this GPU
        assert c.compute(); want = c.draw()
        for f in frames:
            assert np.array_equal(f, want)
        print("OK", len(frames))
    c.close()
""")


@pytest.mark.parametrize("host_frame", [0, 1])
def test_scheduler_world_2(pkg, tmp_path, host_frame):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    session = f"t2_{os.getpid()}_{host_frame}"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", session, str(host_frame)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    assert "OK 5" in outs[0][0]
