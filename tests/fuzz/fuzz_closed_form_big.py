"""CPU: the library's default walk (host build of the device core, both kinds of top grid) on the BENCH map at the headline
frame size -- every 24th row of 3840x2160 -- from many cameras (on the ground, high above it, every 7th on a voxel edge),
one and two lights, against Oracle-B; classification as in fuzz_closed_form.py.

    python tests/fuzz/fuzz_closed_form_big.py [N=1024] [shell|solid] [FIRST_CAMERA LAST_CAMERA]     (about 17 s per camera and light count on 8 cores)
"""
import sys, time, os
import pathlib; R_ = pathlib.Path(__file__).resolve().parents[2]
for p_ in (R_, R_ / "tests", R_ / "tests" / "fuzz"): sys.path.insert(0, str(p_))
import numpy as np, importlib
import fuzz_closed_form as F
O, emu_lib, pkg = F.O, F.emu_lib, F.pkg
S = pkg.scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = sys.argv[2] if len(sys.argv) > 2 else "shell"
cams = range(int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else range(0, 20)
W, H = 3840, 2160
vol = S.terrain_map(n, variant)
h = S.heightfield(n)
rows = np.arange(0, H, 24)
table_full = O.make_ray_table(W, H)
table = table_full[rows].copy()
tot = dict(px=0, ties=0, far=0, bad=0, grids=0)
for cam in cams:
    for nl in (1, 2):
        pos, d = S.make_camera(n, h, cam, need_zero_bias=(cam % 2 == 0))
        pos = np.array(pos, np.float32)
        if cam % 5 == 0: pos[2] = min(n - 2.5, pos[2] + n * 0.3)
        if cam % 7 == 0: pos[:2] = np.floor(pos[:2])          # camera on a voxel edge
        sc = S.Scene(n, vol, W, len(rows), pos, np.array(d, np.float32), S.make_lights(n, nl), max_distance=3 * n)
        t = time.time()
        ref_rgba, ref_aux, _ = O.raycast(sc, table, shadow_lights=nl, canonical_t=True)
        frames = []
        for use_svo in (3, 4):
            rgba, aux = emu_lib.raycast(sc, table, use_svo=use_svo, shadow_lights=nl)
            problems, ties, far = F.classify(ref_rgba, ref_aux, rgba, aux, nl)
            tot["px"] += rgba.shape[0] * rgba.shape[1]; tot["ties"] += ties; tot["far"] += far
            if problems:
                tot["bad"] += 1
                print("MISMATCH", n, variant, cam, nl, use_svo, "; ".join(problems)[:300], flush=True)
            frames.append(rgba)
        dd = np.abs(frames[0].astype(int) - frames[1].astype(int)).max(-1)
        tie = (ref_aux["flags"] & 4) != 0
        if (dd[~tie] > 0).any():
            tot["bad"] += 1; print("GRIDS DIFFER on non-tie pixels", cam, nl, flush=True)
        tot["grids"] += int((dd > 0).sum())
        print(f"cam {cam} lights {nl} pos {pos.tolist()} max steps {int(ref_aux['steps_total'].max())} lit {float(((ref_aux['flags'] & 1) != 0).mean()):.2f} {time.time() - t:.1f}s {tot}", flush=True)
print("done", n, variant, tot)
