"""CPU-side analysis of the octree kernel's traversal on the bench frame (no GPU needed).

Runs tests/host_emu/libvrprof.so: the device core compiled for the host with counters, executed the way a warp
executes it (32 pixels of an 8x4 block in lockstep).  Prints how the cells / DDA steps of the frame split over
the in-cell walks, how many lookups / node loads there are and a crude issue-slot model (calibrated against the
ncu instruction counts in profiles/).  ANALYSIS ONLY; the numbers that count are measured on the GPU.

    python profiles/walk_profile.py [config=c3] [warp_stride=16]
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import oracle_lib as O  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "c3"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = bench.package().scene
emu = ROOT / "tests" / "host_emu"
subprocess.run(["make", "-C", str(emu), "profile"], check=True, capture_output=True)
lib = C.CDLL(str(emu / "libvrprof.so"))

table = {"c1": (64, 1280, 720), "c2": (256, 1920, 1080), "c3": (1024, 3840, 2160), "c4": (4096, 7680, 4320)}
n, w, h = table[config]
lo, hi = S.terrain_columns(n, "shell")
pos, direction = S.make_camera(n, S.heightfield(n), bench.BENCH_CAMERA)
lights = S.make_lights(n, 1)
atlas = S.synthetic_atlas()
rt = O.make_ray_table(w, h)
# issue-slot model: brick fixed/step, axes fixed / per chained add / per fix-up round, merged fixed/step,
# lookup fixed / per pop / per node load, hit block, per-round loop overhead
model = np.array([45, 21, 150, 1.75, 6, 40, 16, 30, 4, 22, 330, 12, 24], np.float32)
if len(sys.argv) > 3:
    model = np.array([float(v) for v in sys.argv[3].split(",")], np.float32)
out = np.zeros(512, np.float64)
fp = C.POINTER(C.c_float)
lo = np.ascontiguousarray(lo, np.int32); hi = np.ascontiguousarray(hi, np.int32)
need = lib.emu_profile(C.c_int(w), C.c_int(h), rt.ctypes.data_as(fp), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                       C.c_int(n), pos.ctypes.data_as(fp), direction.ctypes.data_as(fp), lights.ctypes.data_as(fp),
                       atlas.ctypes.data_as(C.c_void_p), C.c_int(256), C.c_int(256), C.c_int(16), C.c_int(16), C.c_int(3 * n),
                       C.c_int(stride), model.ctypes.data_as(fp), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(512))
assert need > 0, need
NP, NH = 5, 12
i = 0
def take(k):
    global i
    v = out[i:i + k]; i += k
    return v
cells, steps = take(NP), take(NP)
hist = take(NP * NH).reshape(NP, NH); hist_steps = take(NP * NH).reshape(NP, NH)
lookups, pops, loads, hits, replays, chain, fix, rounds, lane_rounds, rays, adds, jumps = take(12)
model_warp, model_lane = take(2)
comp = take(8)
grp_rounds, grp_adds, grp_model, grp_rays = take(8), take(8), take(8), take(8)
mixed_rounds, alt_warp, alt_small, alt_axes = take(4)
names = ["none", "brick", "axes", "merged", "merged after axes fallback"]
print(f"{config} stride {stride}: rays {rays:.0f} (x{stride} = {rays * stride / 1e6:.2f} M), lane-rounds/ray {lane_rounds / rays:.2f}, "
      f"SIMT round efficiency {lane_rounds / (32 * rounds):.3f}")
print(f"per ray: steps {steps.sum() / rays:.1f}, lookups {lookups / rays:.2f}, pops {pops / rays:.2f}, node loads {loads / rays:.2f}, "
      f"hits {hits / rays:.2f}, replays {replays / rays:.4f}, chained adds {chain / rays:.1f} (literal {adds / rays:.1f}, binade jumps {jumps / rays:.2f}), fix-up rounds {fix / rays:.2f}")
for p in range(NP):
    if cells[p] == 0:
        continue
    print(f"  {names[p]:28s} cells/ray {cells[p] / rays:6.2f}  steps/ray {steps[p] / rays:7.1f}  steps/cell {steps[p] / max(cells[p], 1):6.1f}")
    print("      cells by steps<=2^b: " + " ".join(f"{100 * v / cells[p]:4.1f}" for v in hist[p]))
    print("      steps by steps<=2^b: " + " ".join(f"{100 * v / max(steps[p], 1):4.1f}" for v in hist_steps[p]))
print(f"model: warp slots/frame {model_warp * stride / 1e9:.3f} G (ncu smsp__inst_executed), lane slots {model_lane * stride / 1e9:.2f} G, "
      f"thread/inst {model_lane / model_warp:.1f}")
print("  warp-slot share: " + ", ".join(f"{nm} {100 * v / model_warp:.1f}%" for nm, v in zip(["brick", "axes", "merged", "lookup", "hit", "loop"], comp)))
print("by tile row mod 8 (4 screen rows each): warp rounds " + " ".join(f"{v / grp_rounds.mean():.3f}" for v in grp_rounds)
      + " | chain work " + " ".join(f"{v / grp_adds.mean():.3f}" for v in grp_adds) + " | model " + " ".join(f"{v / grp_model.mean():.3f}" for v in grp_model))
print(f"rounds with both a per-axis walk and a small-cell walk in the warp: {100 * mixed_rounds / rounds:.1f} % of {rounds * stride / 1e6:.2f} M warp rounds")
print(f"phase-batched schedule (small-cell lanes iterate until none is left, then one per-axis cell): model {alt_warp * stride / 1e9:.3f} G warp slots "
      f"vs {model_warp * stride / 1e9:.3f} G now ({100 * alt_warp / model_warp:.1f} %), iterations small {alt_small * stride / 1e6:.2f} M + axes {alt_axes * stride / 1e6:.2f} M")
