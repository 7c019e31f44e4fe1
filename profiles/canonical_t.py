"""CPU experiment (no GPU): how far is SURVEY Appendix E's "canonical-t" walk from the reference walk?

The reference advances intersection_t by repeated rounded additions (kernel:559).  Appendix E proposes the closed
form t_axis(k) = t0_axis + k * delta_axis, which a space-skipping traversal can evaluate in O(1) per cell.  This
script runs the device core's DENSE walk compiled for the host twice on the bench frame -- literally, and with the
closed form (VR_CANON_T: 1 = multiply then add, 2 = one fused multiply-add) -- and reports the north_star statistic:
share of pixels with RGBA max-abs-diff > 1, and first-hit voxel / face mismatches inside and outside the oracle's
degenerate flag (VRO_FL_TIE).  The dense walk is used because the closed form makes the result independent of how
the path is enumerated: an octree kernel built on it would produce exactly these frames.

    python profiles/canonical_t.py [config=c3] [row_stride=4]
"""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import emu_lib  # noqa: E402
import oracle_lib as O  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "c3"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 4
emu = ROOT / "tests" / "host_emu"
src = ["emu.cpp", "../../voxel-raycaster_b200/csrc/vr_octree.cpp"]
flags = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared"]
libs = {}
for mode in (0, 1, 2):
    out = Path("/tmp") / f"libvremu_canon{mode}.so"
    subprocess.run(flags + ([f"-DVR_CANON_T={mode}"] if mode else []) + ["-o", str(out)] + src, check=True, cwd=emu)
    libs[mode] = C.CDLL(str(out))

scene = bench.bench_scene(config)
rows = np.arange(0, scene.height, stride)
table = O.make_ray_table(scene.width, scene.height)[rows].copy()
sub = scene.__class__(scene.n, scene.volume, scene.width, len(rows), scene.cam_pos, scene.cam_dir, scene.lights, atlas=scene.atlas,
                      max_distance=scene.max_distance)


def run(lib):
    emu_lib._lib = lib
    lib.emu_raycast.restype = C.c_int
    return emu_lib.raycast(sub, table, use_svo=False)


ref_rgba, ref_aux = run(libs[0])
res = {"config": config, "rows": int(len(rows)), "pixels": int(ref_rgba.shape[0] * ref_rgba.shape[1])}
tie = (ref_aux["flags"] & 4) != 0
res["tie_pixels_share"] = float(tie.mean())
for mode, name in ((1, "mul_add"), (2, "fma")):
    rgba, aux = run(libs[mode])
    d = np.abs(ref_rgba.astype(np.int16) - rgba.astype(np.int16)).max(-1)
    hit_bad = np.any(ref_aux["hit"] != aux["hit"], axis=-1) | (ref_aux["face"] != aux["face"])
    st_bad = ref_aux["status"] != aux["status"]
    res[name] = {
        "rgba_diff_gt0_share": float((d > 0).mean()), "rgba_diff_gt1_share": float((d > 1).mean()),
        "rgba_max_abs_diff": int(d.max()),
        "first_hit_or_face_mismatch_share": float(hit_bad.mean()),
        "first_hit_or_face_mismatch_outside_tie_flag_share": float((hit_bad & ~tie).mean()),
        "status_mismatch_share": float(st_bad.mean()),
        "steps_total_mismatch_share": float((ref_aux["steps_total"] != aux["steps_total"]).mean()),
    }
print(json.dumps(res, indent=1))
