"""Computes the ALGORITHMIC bytes per frame of a bench workload from the oracle's counters
(SURVEY.md 8d / DESIGN.md "byte model"):

    B = P*(16 + 4) + 8*D_svo + 4*T            (SVO mode)
    B = P*(16 + 4) + 1*S_dense + 4*T          (dense mode)

P pixels (16 B ray-table read + 4 B RGBA8 store), D_svo child-descriptor fetches of the canonical
stack-based descent of the reference-format octree, S_dense DDA steps (1 B map load each), T atlas texel
fetches.  Counted on every `--row-stride`-th row and scaled; the result is committed as
profiles/algorithmic_bytes_<config>.json and read by bench.py (the oracle is NOT run inside the timed
benchmark for this).

    python profiles/make_algorithmic_bytes.py --config c3 --row-stride 8
"""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

pkg = importlib.import_module("voxel-raycaster_b200")
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--row-stride", type=int, default=8)
ap.add_argument("--lights", type=int, default=1)
args = ap.parse_args()

if args.config == "c5":
    # BASELINE configs[4]: 64 cameras x 1920x1080 over the 1024^3 SVO (bench.py run_views): the counters of all views added up
    S = pkg.scene
    n, W, H, views = 1024, 1920, 1080, 64
    base = bench.bench_scene("c3")
    h = S.heightfield(n)
    desc, root = pkg.octree_generate(base.volume)
    tot = None
    t0 = time.time()
    for i in range(views):
        pos, direction = S.make_camera(n, h, i)
        sc = S.Scene(n, base.volume, W, H, pos, direction, base.lights, atlas=base.atlas, max_distance=3 * n)
        _, _, cnt = O.raycast(sc, octree=(desc, root), want_aux=False, want_counters=True, count_svo=True, row_stride=args.row_stride)
        tot = cnt if tot is None else {k: tot[k] + v for k, v in cnt.items()}
    print(f"oracle counters of {views} views over 1/{args.row_stride} of the rows in {time.time() - t0:.1f} s")
    k = views * W * H / tot["pixels"]
    P = views * W * H
    out = {"config": "c5", "views": views, "width": W, "height": H, "n": n, "row_stride": args.row_stride, "pixels": P,
           "primary_rays": tot["primary_rays"] * k, "shadow_rays": tot["shadow_rays"] * k, "dda_steps": tot["dda_steps"] * k,
           "texel_fetches": tot["texel_fetches"] * k, "svo_desc_fetches": tot["svo_desc_fetches"] * k,
           "svo_cell_changes": tot["svo_cell_changes"] * k}
    out["bytes_svo"] = P * 20 + 8 * out["svo_desc_fetches"] + 4 * out["texel_fetches"]
    out["bytes_dense"] = P * 20 + 1 * out["dda_steps"] + 4 * out["texel_fetches"]
    out["rays"] = out["primary_rays"] + out["shadow_rays"]
    (ROOT / "profiles" / "algorithmic_bytes_c5.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))
    sys.exit(0)

scene = bench.bench_scene(args.config, lights=args.lights)
t0 = time.time()
if scene.volume is None:
    # c4 (4096^3): no dense map, hence no reference-format buffer; the path of 2^3 descriptors is read off an occupancy tree
    # with 4^3 children per node built from the column tables (vro_scene::tree64; tests/test_oracle.py holds the two
    # counters equal on maps that have both)
    import emu_lib  # noqa: E402

    emu_lib.set_collapse(False)
    nodes, _, levels = emu_lib.tree_from_columns(scene.columns[0], scene.columns[1])
    print(f"occupancy tree from the column tables: {nodes.shape[0]} nodes, {levels} levels in {time.time() - t0:.1f} s")
    t0 = time.time()
    _, aux, cnt = O.raycast(scene, want_counters=True, count_svo=True, row_stride=args.row_stride, shadow_lights=args.lights,
                            tree64=(nodes, levels))
else:
    desc, root = pkg.octree_generate(scene.volume)
    print(f"reference-format octree: {desc.size} descriptors ({desc.nbytes / 1e6:.1f} MB) in {time.time() - t0:.1f} s")
    t0 = time.time()
    _, aux, cnt = O.raycast(scene, octree=(desc, root), want_counters=True, count_svo=True, row_stride=args.row_stride,
                            shadow_lights=args.lights)
print(f"oracle counters over 1/{args.row_stride} of the rows in {time.time() - t0:.1f} s: {cnt}")
k = scene.height * scene.width / cnt["pixels"]
P = scene.width * scene.height
out = {
    "config": args.config, "scene": scene.name, "width": scene.width, "height": scene.height, "n": scene.n,
    "camera_pos": [float(v) for v in scene.cam_pos], "camera_dir": [float(v) for v in scene.cam_dir],
    "max_distance": scene.max_distance, "row_stride": args.row_stride, "sample_pixels": cnt["pixels"],
    "pixels": P,
    "primary_rays": cnt["primary_rays"] * k, "shadow_rays": cnt["shadow_rays"] * k, "reflect_rays": cnt["reflect_rays"] * k,
    "dda_steps": cnt["dda_steps"] * k, "texel_fetches": cnt["texel_fetches"] * k,
    "svo_desc_fetches": cnt["svo_desc_fetches"] * k, "svo_cell_changes": cnt["svo_cell_changes"] * k,
    "tie_pixels": cnt["tie_pixels"] * k,
    "ref_octree_descriptors": int(desc.size) if scene.volume is not None else None,
    "descriptor_path_source": "reference-format octree" if scene.volume is not None else "4^3 occupancy tree built from the column tables (two 2^3 levels per node)",
}
out["bytes_svo"] = P * 20 + 8 * out["svo_desc_fetches"] + 4 * out["texel_fetches"]
out["bytes_dense"] = P * 20 + 1 * out["dda_steps"] + 4 * out["texel_fetches"]
out["rays"] = out["primary_rays"] + out["shadow_rays"] + out["reflect_rays"]
out["shadow_lights"] = args.lights
path = ROOT / "profiles" / (f"algorithmic_bytes_{args.config}.json" if args.lights == 1 else f"algorithmic_bytes_{args.config}_l{args.lights}.json")
path.write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
