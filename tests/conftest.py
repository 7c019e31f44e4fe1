"""pytest configuration: the `gpu` marker, import paths, shared scene helpers."""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present() -> bool:
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a usable CUDA device: the gpu tests are skipped, not failed (the product
    has no CPU path, so vr_init fails there by design).  With a device they run."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the caster has no CPU path (run with -m gpu on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has a hyphen, so import_module)."""
    import __graft_entry__ as entry

    entry.build()          # no-op when libvrcaster.so / the oracle are up to date
    return importlib.import_module("voxel-raycaster_b200")


@pytest.fixture(scope="session")
def oracle(pkg):
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


def c_trunc_div2(v: int) -> int:
    return int(v / 2)


def oracle_bias(oracle_mod, scene, desc, root):
    """((sub_oct_pos - voxel) * resolution) / 2 with C truncation (kernel:354) from the ORACLE's get_oct_vox."""
    v = np.floor(scene.cam_pos).astype(int)
    _, sub, res, _ = oracle_mod.get_oct_vox(desc, root, scene.n, v)
    return [c_trunc_div2((sub[i] - int(v[i])) * res) for i in range(3)]


AUX_FIELDS = ("hit", "face", "status", "flags", "hit_type", "steps_first", "steps_total")


def assert_same_frame(ref_rgba, ref_aux, got_rgba, got_aux, what=""):
    """Bit-exact comparison of RGBA8 and of every integer aux field."""
    for f in AUX_FIELDS:
        bad = np.argwhere(np.any(np.atleast_3d(ref_aux[f] != got_aux[f]), axis=-1))
        assert bad.size == 0, f"{what}: aux field {f} differs at {len(bad)} pixels, first (y,x)={bad[0]}"
    diff = np.abs(ref_rgba.astype(np.int16) - got_rgba.astype(np.int16))
    bad = np.argwhere(diff.max(axis=-1) > 0)
    assert bad.size == 0, f"{what}: RGBA differs at {len(bad)} pixels (max abs diff {diff.max()}), first (y,x)={bad[0]}"
