#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 voxel ray caster.

Metric (BASELINE.json): Mrays/s and ms/frame, 1024^3 SVO @ 3840x2160 + shadows, at 1/2/4/8 B200.
A "step" is one frame: every pixel casts its primary ray and, on a hit, the shadow ray towards light 0,
through the 64-tree kernel.  With N > 1 (torchrun, one rank per GPU) the frame loop is the C library's
multi-GPU scheduler (vr_mgpu_*, run_mgpu below): the octree is broadcast once from rank 0 with NCCL, the
frame is split into the kernel's 32x4-pixel tiles, tile (tx, ty) on rank (tx + ty) mod N, and every rank's
kernel stores its pixels in place into the frame on the root GPU over NVLink; torch.distributed supplies
the barrier around the timed region and the MAX over ranks.  (--gather direct|p2p|nccl: the round-1 Python
pipelines, kept for comparison.)

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference        # the CPU restatement of the reference kernel, all host threads

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
METRIC = "Mrays/s (primary + shadow rays), 1024^3 SVO @ 3840x2160 + shadows"
BAND_ROWS = int(os.environ.get("VR_BAND_ROWS", "8"))   # rows per interleaved band (multiple of the 4-row CTA tile)
BENCH_CAMERA = 9            # make_camera(index): terrain + sky + shadowed pixels
WALK_NAMES = {0: "merged", 1: "per-axis", 2: "closed-form"}
WALK_NOTES = {
    0: "merged in-cell walk, literal additions: bit-identical to the reference restatement on every pixel",
    1: "per-axis in-cell walk, literal additions: identical to the reference restatement except distance_traveled on exact-tie rays (degenerate, ~0.9 % of pixels, RGBA +-1)",
    2: "closed-form crossing times t(k) = fma(k, delta_t, t0) (SURVEY Appendix E): identical to the oracle's restatement of that form except distance_traveled on exact-tie "
       "rays; against the reference walk RGBA8 within +-1 on 99.98 % of the pixels and first hit / face identical outside the degenerate (voxel-edge) rays "
       "(tests/test_gpu_canonical.py, DESIGN.md section 2)",
}


def package():
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    return importlib.import_module("voxel-raycaster_b200")


def bench_scene(config: str = "c3", with_volume: bool = True, lights: int = 1):
    """The named workload.  c3 = BASELINE.json configs[2]: 1024^3 sparse (shell) terrain, 3840x2160,
    one shadow light (light 0 is the only one the reference kernel reads), max_distance 3N."""
    S = package().scene
    if config == "c4":
        # BASELINE configs[3]: 4096^3 deep SVO (12 levels), 7680x4320, 1 shadow light; the volume (64 GiB) is never
        # materialised: the octree is built from the solid z-range of every column
        n = 4096
        lo, hi = S.terrain_columns(n, "shell")
        pos, direction = S.make_camera(n, hi, BENCH_CAMERA)
        return S.Scene(n, None, 7680, 4320, pos, direction, S.make_lights(n, lights), max_distance=3 * n, name="c4-shell",
                       columns=(lo, hi) if with_volume else None)
    if not with_volume:
        # camera / lights / atlas only (ranks > 0 receive the octree by broadcast)
        table = {"c1": (64, 1280, 720), "c2": (256, 1920, 1080), "c3": (1024, 3840, 2160)}
        n, w, h = table[config]
        pos, direction = S.make_camera(n, S.heightfield(n), BENCH_CAMERA)
        return S.Scene(n, None, w, h, pos, direction, S.make_lights(n, lights), max_distance=3 * n, name=f"{config}-shell")
    return S.make_scene(config, camera_index=BENCH_CAMERA, lights=lights)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe: the nvidia-smi counters, read
    through NVML in-process, sub-millisecond per sample, every 5 ms).
    `prepare()` opens NVML, starts the (parked) sampling thread and MUST be called before the barrier that precedes the
    timed region: nvmlInit takes 10+ ms, a Python thread 0.1 ms to start, and
    anything a rank does between that barrier and its first launch skews the ranks against each other -- the other ranks run
    into the scheduler's frame ring and wait, inside their timed regions (measured: with the set-up after the barrier, and
    one sampler process per rank, an 8-GPU run showed per-rank frame times of 0.09 .. 2.6 ms instead of 0.09,
    profiles/r2c_scale8_*; with NVML opened by rank 0 only, rank 1 of 2 lost 11 ms, profiles/r2c_check_c3_n2.json).
    Entering the context only sets the event the thread waits for.  At N > 1 only rank 0 samples, and the clocks under load are also
    sampled over an untimed continuation of the same frames (run_mgpu), because the timed region is a few milliseconds."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    INTERVAL_S = 0.005

    def __init__(self, index: int, enabled: bool = True, delay_first: bool = False) -> None:
        """enabled = False: a no-op (N > 1: only rank 0 samples).  delay_first: the first sample is taken one interval after
        the start instead of at once (N > 1: a region shorter than the interval then sees no query at all)."""
        self.index, self.samples, self._stop, self._t, self._nvml = index, [], threading.Event(), None, None
        self._go = threading.Event()              # set on entering the context: the thread exists (parked) before that
        self.enabled, self.delay_first = enabled, delay_first

    def _open_nvml(self) -> bool:
        try:
            import pynvml as N

            N.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self._h = N.nvmlDeviceGetHandleByIndex(phys)
            self._mx = N.nvmlDeviceGetMaxClockInfo(self._h, N.NVML_CLOCK_SM)
            self._nvml = N
            return True
        except Exception:
            self._nvml = None
            self._tried = True
            return False

    def _run_nvml(self) -> bool:
        N = self._nvml
        if N is None:
            return False
        try:
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            order = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            if self.delay_first and self._stop.wait(self.INTERVAL_S):
                return True
            while True:
                sm = N.nvmlDeviceGetClockInfo(self._h, N.NVML_CLOCK_SM)
                try:
                    reasons = N.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    reasons = N.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append([str(sm), str(self._mx)] + ["Active" if reasons & bits[k] else "Not Active" for k in order])
                if self._stop.wait(self.INTERVAL_S):
                    break
            return True
        except Exception:
            return bool(self.samples)

    def _run(self) -> None:
        self._go.wait()
        if self._stop.is_set():
            return
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def prepare(self):
        """opens NVML (10+ ms) and starts the sampling thread, parked (thread start-up: 0.1 ms): call BEFORE the barrier in
        front of the timed region.  Entering the context then only sets an event."""
        if self.enabled and self._t is None:
            if self._nvml is None and not getattr(self, "_tried", False):
                self._open_nvml()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __enter__(self):
        self.prepare()                                # (not prepared: still works, but skews this rank's start)
        self._go.set()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        self._go.set()
        if self._t is not None:
            self._t.join(timeout=10)

    def summary(self) -> dict:
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def load_algorithmic_bytes(config: str, lights: int = 1) -> dict | None:
    p = ROOT / "profiles" / (f"algorithmic_bytes_{config}.json" if lights == 1 else f"algorithmic_bytes_{config}_l{lights}.json")
    return json.loads(p.read_text()) if p.exists() else None


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def use_all_host_threads() -> int:
    """The CPU legs run on every host core the process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    the OpenMP runtime of the oracle / reference libraries would obey: override it (before those libraries initialise
    OpenMP, and through omp_set_num_threads for a runtime that is already up) and return the team size OpenMP will
    actually use -- that number, not the core count, is what the JSON line reports as `cores`."""
    import ctypes as C

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        g = C.CDLL("libgomp.so.1")
        g.omp_set_num_threads(C.c_int(n))
        g.omp_get_max_threads.restype = C.c_int
        return int(g.omp_get_max_threads())
    except OSError:
        return n


def oracle_sample(scene, row_stride: int, threads: int = 0, lights: int = 1) -> tuple[float, int, int]:
    """Times the CPU restatement of the reference kernel (dense DDA over the char map, kernel:555-570) on
    every row_stride-th row.  Returns (seconds, rays in the sample, host threads)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O

    table = O.make_ray_table(scene.width, scene.height)
    O.raycast(scene, table, rows=(0, 8), want_aux=False)                      # warm the library / page in
    t0 = time.perf_counter()
    _, aux, cnt = O.raycast(scene, table, want_aux=True, row_stride=row_stride, threads=threads, shadow_lights=lights,
                             want_counters=lights > 1)
    dt = time.perf_counter() - t0
    a = aux[::row_stride]
    rays = int((a["status"] != O.ST_SKIP_PRIMARY).sum() + ((a["flags"] & O.FL_LIT) != 0).sum())
    if lights > 1:
        rays = int(cnt["primary_rays"] + cnt["shadow_rays"])
    return dt, rays, threads or use_all_host_threads()


class CpuReference:
    """The reference's CPU implementation of the path on every row_stride-th row of the frame, all host threads.
    kind "reference": the reference's OWN kernel source (kernels/ray_caster_kernel.cl) compiled for the CPU by g++ through
    oracle/ref_shim/cl_shim.h (oracle/_ref, built where /root/reference exists; tests/test_reference_kernel.py) -- the
    `md` build (max_distance lifted to the workload's) for maps up to 256^3, the `wide` build (additionally the kernel's
    8-entry private stacks widened to 32, kernel:119-124) beyond.  kind "port": the oracle's restatement, when that
    library is absent, the scene has no dense map (4096^3) or more than one light is requested (extension)."""

    def __init__(self, scene, row_stride: int, lights: int = 1) -> None:
        self.scene, self.stride, self.lights = scene, row_stride, lights
        self.kind, self.lib = "port", None
        use_all_host_threads()
        self.note = "reference OpenCL kernel restated in C++ (oracle/, no OpenCL runtime in the image)"
        self.dt0, self.rays, self.threads = oracle_sample(scene, row_stride, lights=lights)      # also the ray count of the sample
        if lights == 1 and scene.volume is not None:
            try:
                sys.path.insert(0, str(ROOT / "tests"))
                import ref_kernel_lib as R

                build = True if scene.n <= 256 else "wide"
                if R.available(build):
                    R.lib(build)                                      # loads (or raises) here, not inside the timed call
                    octree = package().octree_generate(scene.volume)
                    self.lib, self.build, self.kind, self.octree = R, build, "reference", octree
                    self.note = ("kernels/ray_caster_kernel.cl itself, compiled for the CPU by g++ through oracle/ref_shim/cl_shim.h (OpenMP over rows); "
                                 "max_distance lifted to the workload's" + ("" if build is True else ", private stacks widened 8 -> 32 entries"))
            except Exception as e:                                    # a missing / unloadable library must not cost the bench line
                sys.stderr.write(f"bench.py: reference kernel library unusable ({e}); timing the port instead\n")
                self.lib, self.kind = None, "port"

    def run(self) -> float:
        """seconds for one pass over the sample"""
        if self.lib is not None:
            try:
                t0 = time.perf_counter()
                self.lib.raycast(self.scene, octree=self.octree, lifted=self.build, row_stride=self.stride)
                return time.perf_counter() - t0
            except Exception as e:
                sys.stderr.write(f"bench.py: reference kernel run failed ({e}); timing the port instead\n")
                self.lib, self.kind = None, "port"
                self.note = "reference OpenCL kernel restated in C++ (oracle/, no OpenCL runtime in the image)"
        return oracle_sample(self.scene, self.stride, lights=self.lights)[0]


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path (see CpuReference) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()                       # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core
    scene = bench_scene(args.config, lights=args.lights)
    stride = args.ref_row_stride
    cpu = CpuReference(scene, stride, args.lights)
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu.run()
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    rays, threads = cpu.rays, cpu.threads
    value = rays / (ms / 1e3) / 1e6
    sample = f"every {stride}th row of the {scene.width}x{scene.height} frame ({rays} rays per step)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"{args.config}: {scene.n}^3 shell terrain, {scene.width}x{scene.height}, {args.lights} shadow light{'s' if args.lights > 1 else ''}, dense DDA on CPU",
                   "note": cpu.note + "; ms_per_step is for the sample"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def c5_roofline(ms_per_batch: float, world: int) -> dict | None:
    """algorithmic bytes of the 64-view batch (profiles/algorithmic_bytes_c5.json, oracle counters of all views) over the
    batch time: every launch renders one view, a rank renders 64 / world of them back to back"""
    ab = load_algorithmic_bytes("c5")
    if not ab:
        return None
    peak, peak_how = measured_peak_gbs()
    per_view = float(ab["bytes_svo"]) / ab["views"]
    achieved = float(ab["bytes_svo"]) / world / (ms_per_batch / 1e3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_how,
            "kernel": "vr_svo_kernel", "kernel_ms": ms_per_batch / (ab["views"] / world), "algorithmic_bytes_per_launch": per_view,
            "bytes_model": "P*(16+4) + 8*D_svo + 4*T summed over the 64 views (profiles/algorithmic_bytes_c5.json); one launch = one view"}


def run_views(args, pkg, torch, dist, rank, world, local_rank, dev) -> None:
    """--config c5 (BASELINE configs[4], not the headline): 64 cameras x 1920x1080 over the 1024^3 SVO, view-batch
    split: view v is rendered by rank v % world with vr_compute_views; a step is the whole batch."""
    S = pkg.scene
    n, W, H, views = 1024, 1920, 1080, 64
    h = S.heightfield(n)
    cams = np.array([np.concatenate([d, p]) for p, d in (S.make_camera(n, h, i) for i in range(views))], dtype=np.float32)
    mine = cams[rank::world].copy()
    c = pkg.CUDACaster()

    def must(ok, what):
        if not ok:
            raise RuntimeError(f"{what}: {c.last_error()}")

    must(c.init(local_rank), "init")
    must(c.add_to_settings_buffer("octree_dimensions", "OCTDIM", n), "OCTDIM")
    must(c.add_to_settings_buffer("using_octree", "OCTENABLED", 0), "OCTENABLED")
    must(c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", 3 * n), "MAX_DISTANCE")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    must(c.set_stream(stream.cuda_stream), "set_stream")
    lo, hi = S.terrain_columns(n, "shell")
    must(c.assign_columns(lo, hi), "assign_columns")        # every rank builds the 3.9 MB octree itself (0.2 s)
    cam_dir, cam_pos = mine[0, :2].copy(), mine[0, 2:].copy()
    must(c.assign_camera(cam_dir, cam_pos), "assign_camera")
    must(c.create_viewport(W, H, 56.25, 90.0), "create_viewport")
    lights = S.make_lights(n, 1)
    must(c.assign_lights(lights), "assign_lights")
    must(c.create_texture_atlas(S.synthetic_atlas(), (16, 16)), "atlas")
    must(c.validate(), "validate")
    must(c.set_option("walk", args.walk), "walk")
    must(c.set_option("directed_grid", args.directed_grid), "directed_grid")
    frames = torch.empty((len(mine), H, W, 4), dtype=torch.uint8, device=dev)
    # rays per batch from one untimed aux pass per view
    must(c.enable_aux(True), "aux")
    rays = 0
    for cam in mine:
        cam_dir[:], cam_pos[:] = cam[:2], cam[2:]
        must(c.compute_into(frames[0].data_ptr()), "compute_into")
        torch.cuda.synchronize()
        aux = c.read_aux()
        rays += int((aux["status"] != 0).sum() + ((aux["flags"] & 1) != 0).sum())
    must(c.enable_aux(False), "aux off")
    total = torch.tensor([rays], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total)
    rays = int(total.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        must(c.compute_views(mine, frames.data_ptr()), "compute_views")
    sampler = ClockSampler(local_rank, enabled=(rank == 0), delay_first=(world > 1)).prepare()      # NVML opened before the barrier
    barrier()
    l0 = c.stats().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            must(c.compute_views(mine, frames.data_ptr()), "compute_views")
        ev1.record(stream)
        barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    launches = c.stats().kernel_launches - l0
    # e2e: every frame of the batch is copied to pinned host memory inside the timed region
    host = torch.empty((len(mine), H, W, 4), dtype=torch.uint8).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        must(c.compute_views(mine, frames.data_ptr()), "compute_views")
        host.copy_(frames, non_blocking=True)
    barrier()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(e2e.item()) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "Mrays/s (primary + shadow rays), 64 views x 1920x1080 over 1024^3 SVO", "value": rays / (ms / 1e3) / 1e6,
            "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gpu_launches": int(launches),
            "config": {"workload": "c5: 64 random cameras x 1920x1080, 1024^3 shell terrain SVO, primary + 1 shadow light (BASELINE configs[4], not the headline)",
                       "parallelism": f"views{world}: view v on rank v % {world}", "views": views, "rays_per_batch": rays,
                       "walk": WALK_NAMES[args.walk], "top_grid": "directed" if args.directed_grid else "undirected",
                       "ms_per_view": ms / (views / world) if world else None},
            "roofline": c5_roofline(ms, world), "cpu_baseline": None,
            "e2e": {"value": rays / (e2e_ms / 1e3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(mine.nbytes), "d2h_bytes_per_step": int(frames.numel())},
            "clocks": clocks.summary()}))
    c.close()
    if world > 1:
        dist.destroy_process_group()


def run_mgpu(args, pkg, torch, dist, rank, world, local_rank, dev) -> None:
    """N > 1 with the multi-GPU frame scheduler of the C library (voxel-raycaster_b200/csrc/vr_mgpu.cu: vr_mgpu_*): one
    process per GPU, the octree broadcast once with NCCL from rank 0, every rank's kernel storing its 32x4-pixel tiles
    in place into the frame on the root GPU over NVLink, completion through device-stored counters polled by the root's
    CPU (no per-frame collective), three frame buffers.  torch.distributed only supplies the barrier around the timed
    region and the MAX over ranks of the device times, as the bench contract asks."""
    import ctypes as C

    use_svo = True
    scene = bench_scene(args.config, with_volume=(rank == 0), lights=args.lights)
    c = pkg.CUDACaster()

    def must(ok, what):
        if not ok:
            raise RuntimeError(f"{what}: {c.last_error()}")

    must(c.init(local_rank), "init")
    must(c.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n), "OCTDIM")
    must(c.add_to_settings_buffer("using_octree", "OCTENABLED", 0), "OCTENABLED")
    must(c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance), "MAX_DISTANCE")
    if args.lights > 1:
        must(c.add_to_settings_buffer("light_count", "LIGHT_COUNT", args.lights), "LIGHT_COUNT")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    must(c.set_stream(stream.cuda_stream), "set_stream")
    t_build = time.perf_counter()
    if rank == 0:
        if scene.volume is not None:
            must(c.assign_map(scene.volume), "assign_map")
        else:
            must(c.assign_columns(scene.columns[0], scene.columns[1]), "assign_columns")
    t_build = time.perf_counter() - t_build
    must(c.assign_camera(scene.cam_dir, scene.cam_pos), "assign_camera")
    W, H = scene.width, scene.height
    must(c.create_viewport(W, H, 0.625 * 90.0, 90.0), "create_viewport")
    must(c.assign_lights(scene.lights), "assign_lights")
    must(c.create_texture_atlas(scene.atlas, (scene.tile, scene.tile)), "create_texture_atlas")
    # a session name unique to this launch, the same on every rank
    nonce = [f"{os.environ.get('MASTER_PORT', '0')}_{os.getpid()}"]
    dist.broadcast_object_list(nonce, src=0)
    session = f"b{nonce[0]}"
    must(c.mgpu_init(session + "d", world, rank, 0), "mgpu_init")
    must(c.mgpu_broadcast_octree(), "mgpu_broadcast_octree")
    must(c.validate(), "validate")
    must(c.set_option("walk", args.walk), "walk")
    must(c.set_option("directed_grid", args.directed_grid), "directed_grid")
    st0 = c.stats()
    bcast_bytes = int(st0.native_bytes)

    # rays per frame: one untimed full-frame pass with aux records on rank 0
    primary = shadow = node_fetches = lookups = steps_total = 0
    if rank == 0:
        must(c.set_tiles(1, 0) and c.enable_aux(True), "aux")
        tmp = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        must(c.compute_into(tmp.data_ptr()), "compute_into")
        torch.cuda.synchronize()
        aux = c.read_aux()
        primary, shadow = int((aux["status"] != 0).sum()), int(((aux["flags"] & 1) != 0).sum()) * args.lights
        node_fetches, lookups = int(aux["node_fetches"].astype(np.int64).sum()), int(aux["lookups"].astype(np.int64).sum())
        steps_total = int(aux["steps_total"].astype(np.int64).sum())
        must(c.enable_aux(False) and c.set_tiles(world, rank), "aux off")
        del aux, tmp
    rays = primary + shadow

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def frames(count):
        """`count` frames; the root collects (waits for and releases) frame k - 2 while frames k - 1 and k render, so the
        CPU stays two frames ahead of the GPUs; the other ranks never wait -- the frame ring of the scheduler (4 buffers)
        holds them back when they run ahead of the root"""
        issued = []
        for i in range(count):
            k = c.mgpu_frame()
            must(k >= 0, "mgpu_frame")
            issued.append(k)
            if rank == 0 and len(issued) > 2:
                done = issued.pop(0)
                must(c.mgpu_frame_wait(done) is not None, "mgpu_frame_wait")
                c.mgpu_frame_release(done)
        must(c.mgpu_flush(), "mgpu_flush")               # `stream` now waits for every frame: events on it bracket them all
        ptr = 0
        for k in issued:                                  # the last frames: everybody waits (the root for all ranks)
            ptr = c.mgpu_frame_wait(k)
            must(ptr is not None, "mgpu_frame_wait")
            if k != issued[-1]:
                c.mgpu_frame_release(k)
        return issued[-1], ptr

    last, ptr = frames(args.warmup)
    c.mgpu_frame_release(last)
    # everything slow happens BEFORE the barrier: what a rank does between the barrier and its first launch makes the other
    # ranks wait for it at the scheduler's frame ring, inside their timed regions (ClockSampler)
    sampler = ClockSampler(local_rank, enabled=(rank == 0), delay_first=True).prepare()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = c.stats().kernel_launches
    barrier()
    # the ranks leave the NCCL barrier + device synchronize some tens of microseconds apart; the scheduler's own CPU
    # rendezvous on the shared segment aligns them to about a microsecond (best effort: its result is not checked --
    # a timed region of 20 frames at 8 GPUs is under 2 ms, and a rank that starts late costs the others that time)
    c.mgpu_barrier()
    t_start_ns = time.clock_gettime_ns(time.CLOCK_MONOTONIC)      # (system-wide clock: comparable between the ranks of a node)
    with sampler as clocks:
        ev0.record(stream)
        last, ptr = frames(args.steps)
        ev1.record(stream)
        barrier()
    launches = c.stats().kernel_launches - launches0
    device_checksum = None
    if rank == 0:
        frame = torch.as_tensor(pkg.tiles._RawCuda(ptr, (H, W, 4)), device=dev)
        device_checksum = int(frame[::64, ::64].sum().item())
    c.mgpu_frame_release(last)
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(per_rank, t_ms)                      # how evenly the tile interleave spreads the frame
    per_rank_ms = [round(float(t.item()) / args.steps, 4) for t in per_rank]
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    # clocks under load: the timed region is a few milliseconds (one or two 5 ms samples at best), so rank 0 also samples
    # over an untimed continuation of the same frames (>= 60 ms of them)
    # That continuation is timed as well (config.sustained: the same frames back to back, hundreds instead of K, so that the
    # fill and drain of the frame pipeline and the start skew of the ranks do not weigh; rank 0 polls NVML during it).
    load_frames = max(args.steps, int(60.0 / max(ms_per_step, 1e-3)))
    load_sampler = ClockSampler(local_rank, enabled=(rank == 0)).prepare()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c.mgpu_barrier()
    with load_sampler as load_clocks:
        sv0.record(stream)
        last, ptr = frames(load_frames)
        sv1.record(stream)
        barrier()
    c.mgpu_frame_release(last)
    clocks.samples += load_clocks.samples
    sus_ms = torch.tensor([sv0.elapsed_time(sv1)], dtype=torch.float64, device=dev)
    dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
    sustained = {"frames": load_frames, "ms_per_frame": float(sus_ms.item()) / load_frames,
                 "how": "untimed continuation of the timed frames, CUDA events, MAX over ranks; rank 0 samples its clocks every 5 ms during it",
                 "clocks": load_clocks.summary()}
    # how far apart the ranks entered the timed region (CLOCK_MONOTONIC right after the alignment)
    starts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(starts, torch.tensor([t_start_ns], dtype=torch.int64, device=dev))
    starts = [int(t.item()) for t in starts]
    start_skew_us = (max(starts) - min(starts)) / 1e3
    must(c.mgpu_shutdown(), "mgpu_shutdown")

    # ---- end to end with the result in HOST memory: every rank copies its bands into a shared pinned host frame
    must(c.mgpu_init(session + "h", world, rank, c.MGPU_HOST_FRAME), "mgpu_init host")
    last, ptr = frames(args.warmup)
    c.mgpu_frame_release(last)
    barrier()
    c.mgpu_barrier()
    t0 = time.perf_counter()
    last, ptr = frames(args.steps)                      # returns on the root when every rank's bands of the last frame are in host memory
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)      # per rank: aligned start -> its frames done; MAX = the root's
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(e2e_s.item()) / args.steps
    checksum = 0
    if rank == 0:
        host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(H, W, 4))
        checksum = int(host[::64, ::64].astype(np.int64).sum())
    c.mgpu_frame_release(last)
    must(c.mgpu_shutdown(), "mgpu_shutdown host")

    if rank == 0:
        value = rays / (ms_per_step / 1e3) / 1e6
        peak, peak_how = measured_peak_gbs()
        ab = load_algorithmic_bytes(args.config, args.lights)
        algo_bytes = float(ab["bytes_svo"]) if ab else None
        st = c.stats()
        roofline = None
        if algo_bytes:
            # consecutive frames overlap on the stream of a rank, so the effective duration per launch is the step time
            achieved = algo_bytes / world / (ms_per_step / 1e3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "peak_source": peak_how, "kernel": "vr_svo_kernel", "kernel_ms": ms_per_step,
                        "algorithmic_bytes_per_launch": algo_bytes / world,
                        "bytes_model": "P*(16+4) + 8*D_svo + 4*T from oracle counters (profiles/algorithmic_bytes_%s.json), 1/N per rank" % args.config}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": int(launches),
            "config": {"workload": f"{args.config}: {scene.n}^3 shell terrain SVO, {W}x{H}, primary + {args.lights} shadow light{'s (multi-light extension)' if args.lights > 1 else ''}, max_distance {scene.max_distance}",
                       "mode": "svo", "walk": WALK_NOTES[args.walk], "top_grid": "directed" if args.directed_grid else "undirected",
                       "kernel_variant": "static 32x4 tiles, 128-thread CTAs, 8 CTAs/SM",
                       "parallelism": f"tiles{world}: C-library scheduler (vr_mgpu_*), 2-D interleave of 32x4-pixel tiles ((tx + ty) % {world}), every rank's kernel stores its pixels in place "
                                      "into the root GPU's frame over NVLink (CUDA IPC), device-stored completion counters polled by the root's CPU, 4 frame buffers; "
                                      "1 launch per frame and rank (the completion counter is a stream memory operation, cuStreamWriteValue64)",
                       "l2": "per-frame streams (ray table + image) exceed the 126 MB L2 at N = 1; the octree stays L2-resident by design",
                       "per_rank_ms_per_frame": per_rank_ms, "start_skew_us": round(start_skew_us, 1), "sustained": sustained,
                       "clock_sampling": f"rank 0 only, 5 ms interval, first query 5 ms into the timed region; plus a continuation of {load_frames} "
                                         "frames right after it (config.sustained; the timed region of a multi-GPU run is a few milliseconds)",
                       "rays_per_frame": rays, "primary_rays": primary, "shadow_rays": shadow, "primary_mpix_per_s": primary / (ms_per_step / 1e3) / 1e6,
                       "tree_nodes": int(st.native_nodes), "tree_bytes": int(st.native_bytes), "levels": int(st.levels),
                       "octree_broadcast_bytes": bcast_bytes, "octree_broadcast": "ncclBroadcast from rank 0 (vr_mgpu_broadcast_octree)", "scene_build_s": round(t_build, 2),
                       "dda_steps_per_frame": steps_total, "octree_lookups_per_frame": lookups, "frame_checksum": checksum, "device_frame_checksum": device_checksum},
            "roofline": roofline, "cpu_baseline": None,
            "e2e": {"value": rays / (e2e_ms / 1e3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": (5 * 4 + 10 * 4 + 64 * 8) * world,
                    "d2h_bytes_per_step": W * H * 4,
                    "how": "vr_mgpu_* with VR_MGPU_HOST_FRAME: every rank copies its row bands D2H into a shared page-locked host frame (POSIX shm), all PCIe links in parallel"},
            "clocks": clocks.summary()}))
    c.close()
    dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--mode", default="svo", choices=["svo", "dense"])
    ap.add_argument("--cpu-row-stride", type=int, default=2, help="oracle sample for cpu_baseline (every Nth row; ~25 core-seconds at c3)")
    ap.add_argument("--ref-row-stride", type=int, default=4, help="oracle sample per step for --impl reference (every Nth row)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lights", type=int, default=1, help="shadow lights: 1 = the reference kernel (light 0 only, the headline); 2 = BASELINE configs[2]'s wording, "
                    "through the multi-light extension (one shadow ray per light from the same hit)")
    ap.add_argument("--persistent", type=int, default=0, help="1 = persistent-warp octree kernel")
    ap.add_argument("--refill-min", type=int, default=8)
    ap.add_argument("--ctas-per-sm", type=int, default=3)
    ap.add_argument("--l2-persist", type=int, default=0, help="1 = cudaAccessPolicyWindow over the octree nodes")
    ap.add_argument("--walk", type=int, default=2, choices=[0, 1, 2],
                    help="in-cell walk of the octree kernel: 2 = closed-form crossing times (default: within BASELINE.json's tolerance of the reference walk, "
                    "see DESIGN.md section 2), 1 = literal additions walked per axis (exact except the step count of exact-tie rays), "
                    "0 = literal additions, merged walk (bit-identical to the reference on every pixel)")
    ap.add_argument("--directed-grid", type=int, default=1, choices=[0, 1],
                    help="top grid of the closed-form walk: 1 = one table per direction octant of the ray (default), 0 = the single undirected table")
    ap.add_argument("--overlap-frames", type=int, default=1, help="N > 1: 1 = consecutive frames are launched on two alternating streams (the next frame's "
                    "first CTAs fill the tail of the current one); the kernel events then overlap, so roofline.kernel_ms is the step time")
    ap.add_argument("--host-frame", default="shared", choices=["shared", "root"],
                    help="N > 1 end-to-end leg: 'shared' = every rank copies its bands into a shared pinned host frame; 'root' = rank 0 copies the gathered frame")
    ap.add_argument("--gather", default="mgpu", choices=["mgpu", "direct", "p2p", "nccl"],
                    help="N > 1 frame assembly on the root GPU: 'mgpu' (default) = the C library's scheduler (vr_mgpu_*: tiles stored in place over NVLink, "
                    "device-stored completion counters); the others are the round-1 Python pipelines kept for comparison: 'direct' = the same tile stores with an "
                    "NCCL all_reduce as completion token, 'p2p' = row bands + copy-engine push into the root's frame, 'nccl' = row bands + all_gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly one JSON line: whatever NCCL logs (NCCL_DEBUG=VERSION/INFO prints to stdout) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the caster has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pkg = package()
    if args.config == "c5":
        run_views(args, pkg, torch, dist, rank, world, local_rank, dev)
        return
    if world > 1 and args.gather == "mgpu" and args.mode == "svo":
        run_mgpu(args, pkg, torch, dist, rank, world, local_rank, dev)
        return
    use_svo = args.mode == "svo"
    scene = bench_scene(args.config, with_volume=(rank == 0 or not use_svo), lights=args.lights)
    c = pkg.CUDACaster()

    def must(ok, what):
        if not ok:
            raise RuntimeError(f"{what}: {c.last_error()}")

    must(c.init(local_rank), "init")
    must(c.add_to_settings_buffer("octree_dimensions", "OCTDIM", scene.n), "OCTDIM")
    must(c.add_to_settings_buffer("using_octree", "OCTENABLED", 0 if use_svo else 1), "OCTENABLED")
    must(c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", scene.max_distance), "MAX_DISTANCE")
    if args.lights > 1:                          # multi-light extension (the reference kernel reads light 0 only)
        must(c.add_to_settings_buffer("light_count", "LIGHT_COUNT", args.lights), "LIGHT_COUNT")
    stream = torch.cuda.Stream(device=dev)       # a real (non-default) stream shared by torch, NCCL and the caster
    torch.cuda.set_stream(stream)
    must(c.set_stream(stream.cuda_stream), "set_stream")
    t_build = time.perf_counter()
    if scene.volume is not None:
        must(c.assign_map(scene.volume), "assign_map")          # dense upload + 64-tree build
    elif scene.columns is not None:
        must(c.assign_columns(scene.columns[0], scene.columns[1]), "assign_columns")   # 64-tree from column z-ranges
    t_build = time.perf_counter() - t_build
    bcast_bytes = 0
    if world > 1 and use_svo:
        # "the octree is broadcast once": rank 0's 64-tree -> every rank, NCCL over NVLink
        meta = torch.zeros(4, dtype=torch.int64, device=dev)
        if rank == 0:
            meta[:] = torch.tensor(c.native_tree_info(), dtype=torch.int64)
        dist.broadcast(meta, 0)
        nb, tb, levels, dim = [int(v) for v in meta.tolist()]
        nodes = torch.empty(nb, dtype=torch.uint8, device=dev)
        types = torch.empty(tb, dtype=torch.uint8, device=dev)
        if rank == 0:
            must(c.native_tree_copy(nodes.data_ptr(), types.data_ptr()), "native_tree_copy")
        dist.broadcast(nodes, 0)
        dist.broadcast(types, 0)
        torch.cuda.synchronize()
        if rank != 0:
            must(c.assign_native_tree(nodes.data_ptr(), nb, types.data_ptr(), tb, levels, dim), "assign_native_tree")
        bcast_bytes = nb + tb
        del nodes, types
    must(c.assign_camera(scene.cam_dir, scene.cam_pos), "assign_camera")
    must(c.set_bands(BAND_ROWS, world, rank), "set_bands")
    must(c.create_viewport(scene.width, scene.height, 0.625 * 90.0, 90.0), "create_viewport")
    must(c.assign_lights(scene.lights), "assign_lights")
    must(c.create_texture_atlas(scene.atlas, (scene.tile, scene.tile)), "create_texture_atlas")
    must(c.validate(), "validate")
    must(c.set_option("persistent", args.persistent) and c.set_option("refill_min", args.refill_min)
         and c.set_option("ctas_per_sm", args.ctas_per_sm), "set_option")
    must(c.set_option("walk", args.walk), "walk")
    must(c.set_option("directed_grid", args.directed_grid), "directed_grid")
    if args.l2_persist and use_svo:
        must(c.set_option("l2_persist", 1), "l2_persist")

    W, H = scene.width, scene.height
    layout = pkg.tiles.BandLayout(H, W, BAND_ROWS, world)
    max_bands = layout.max_bands
    slab = torch.full((layout.slab_rows, W, 4), 0, dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * layout.slab_rows, W, 4), dtype=torch.uint8, device=dev) if world > 1 else None
    frame = torch.empty((max_bands * world * BAND_ROWS, W, 4), dtype=torch.uint8, device=dev) if (world > 1 and rank == 0) else None
    host_frame = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() if rank == 0 else None

    def render_step() -> None:
        """device-resident step: render this rank's bands, gather the slabs on rank 0, un-interleave."""
        must(c.compute_into(slab.data_ptr()), "compute_into")
        if world > 1:
            pkg.tiles.gather_frame(layout, slab, gathered, frame, dist, rank)

    # rays per frame, counted on the device from the aux records of one untimed frame
    must(c.enable_aux(True), "enable_aux")
    render_step()
    torch.cuda.synchronize()
    aux = c.read_aux()[: layout.local_rows(rank)]
    counts = torch.tensor([int((aux["status"] != 0).sum()), int(((aux["flags"] & 1) != 0).sum()),
                           int(aux["node_fetches"].astype(np.int64).sum()), int(aux["lookups"].astype(np.int64).sum()),
                           int(aux["steps_total"].astype(np.int64).sum())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    primary, shadow, node_fetches, lookups, steps_total = [int(v) for v in counts.tolist()]
    shadow *= args.lights                        # one shadow ray per light from every lit hit
    rays = primary + shadow
    must(c.enable_aux(False), "enable_aux off")
    del aux

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps bracketed by barrier + synchronize, CUDA events on the launch stream.
    # N > 1: the gather of frame k overlaps the rendering of frame k+1 (tiles.FramePipeline, double buffered).
    direct = world > 1 and args.gather == "direct"
    if direct:
        # 2-D tile interleave, rendered in place into the root's frame over NVLink: no slab, no gather
        must(c.set_bands(BAND_ROWS, 1, 0) and c.set_tiles(world, rank), "set_tiles")
    pipe = pkg.tiles.FramePipeline(layout, dev, dist, rank, lambda ptr: must(c.compute_into(ptr), "compute_into"),
                                   caster=c if args.gather in ("p2p", "direct") else None, direct=direct) if world > 1 else None
    overlap = pipe is not None and args.overlap_frames == 1
    if overlap:
        # consecutive frames on two streams: the first CTAs of frame k+1 fill the SMs the tail of frame k leaves idle
        stream_b = torch.cuda.Stream(device=dev)
        pipe.alternate_streams([stream, stream_b], lambda s: must(c.set_stream(s.cuda_stream), "set_stream"))
    for _ in range(args.warmup):
        pipe.step() if pipe else render_step()
    if pipe:
        pipe.drain()
    sampler = ClockSampler(local_rank, enabled=(rank == 0), delay_first=(world > 1)).prepare()      # NVML opened before the barrier
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = c.stats().kernel_launches
    with sampler as clocks:
        ev0.record(stream)
        for i in range(args.steps):
            if pipe:
                def timed_render(ptr, i=i):               # events hug the kernel, not the wait for a free slab
                    ks = pipe.render_streams[pipe.k & 1] if pipe.render_streams is not None else stream
                    kev[i][0].record(ks)
                    must(c.compute_into(ptr), "compute_into")
                    kev[i][1].record(ks)
                pipe.render = timed_render
                pipe.step()
            else:
                kev[i][0].record(stream)
                must(c.compute_into(slab.data_ptr()), "compute_into")
                kev[i][1].record(stream)
        if pipe:
            pipe.render = lambda ptr: must(c.compute_into(ptr), "compute_into")
            pipe.drain()
        ev1.record(stream)
        barrier()
    launches = c.stats().kernel_launches - launches0
    device_checksum = None
    if overlap:
        pipe.render_streams = None
        must(c.set_stream(stream.cuda_stream), "set_stream")
    if pipe is not None and rank == 0:
        torch.cuda.synchronize()
        device_checksum = int(pipe.frame[: scene.height][::64, ::64].sum().item())     # the frame assembled on the root GPU
    if direct:
        pipe.direct = False                          # the legs below (other walk, end to end) use row bands again
        pipe.slabs = [torch.zeros((layout.slab_rows, scene.width, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
        pipe.nbuf = 2
        must(c.set_tiles(1, 0) and c.set_bands(BAND_ROWS, world, rank), "set_bands")
    t_ms = torch.tensor([ev0.elapsed_time(ev1), float(np.mean([a.elapsed_time(b) for a, b in kev]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = [float(v) for v in t_ms.tolist()]
    ms_per_step = total_ms / args.steps
    if overlap:
        kernel_ms = ms_per_step      # launches of consecutive frames overlap: effective duration per launch
    value = rays / (ms_per_step / 1e3) / 1e6

    # a longer look at the same thing (N = 1): back-to-back frames for at least half a second, with their own clock samples
    sustained = None
    if world == 1:
        n_sus = max(args.steps, int(0.6e3 / max(ms_per_step, 1e-3)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank, enabled=(rank == 0)).prepare() as sus_clocks:
            s0.record(stream)
            for _ in range(n_sus):
                must(c.compute_into(slab.data_ptr()), "compute_into")
            s1.record(stream)
            torch.cuda.synchronize()
        sustained = {"frames": n_sus, "ms_per_frame": s0.elapsed_time(s1) / n_sus, "clocks": sus_clocks.summary()}

    # the other in-cell walks, for the record (a few frames outside the timed region, kernel only)
    other_walk_ms = None
    if use_svo and world == 1:
        other_walk_ms = {}
        for w in (0, 1, 2):
            if w == args.walk:
                continue
            must(c.set_option("walk", w), "walk")
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(3 + 10):
                if i == 3:
                    o0.record(stream)
                must(c.compute_into(slab.data_ptr()), "compute_into")
            o1.record(stream)
            torch.cuda.synchronize()
            other_walk_ms[WALK_NAMES[w]] = o0.elapsed_time(o1) / 10
        must(c.set_option("walk", args.walk), "walk")
        if args.walk == 2:
            # the same walk over the other kind of top grid (the tables are rebuilt, outside the timed frames)
            must(c.set_option("directed_grid", 0 if args.directed_grid else 1), "directed_grid")
            must(c.compute_into(slab.data_ptr()), "compute_into")
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(3 + 10):
                if i == 3:
                    o0.record(stream)
                must(c.compute_into(slab.data_ptr()), "compute_into")
            o1.record(stream)
            torch.cuda.synchronize()
            other_walk_ms["closed-form, " + ("undirected" if args.directed_grid else "directed") + " top grid"] = o0.elapsed_time(o1) / 10
            must(c.set_option("directed_grid", args.directed_grid), "directed_grid")
            must(c.compute_into(slab.data_ptr()), "compute_into")

    # ---- end to end through the public API with HOST buffers: camera/lights are read from host memory at every
    # call, the frame is copied back to pinned host memory inside the timed region (double buffered at N = 1)
    barrier()
    t0 = time.perf_counter()
    t_done = None
    e2e_how = "frame_begin/frame_end, double buffered D2H on a second stream" if world == 1 else "rank 0 copies every gathered frame D2H"
    if world == 1:
        c.set_bands(BAND_ROWS, 1, 0)
        must(c.frame_begin(), "frame_begin")
        for i in range(args.steps - 1):
            must(c.frame_begin(), "frame_begin")
            host = c.frame_end()
        host = c.frame_end()
        checksum = int(host[::64, ::64].astype(np.int64).sum())
    elif args.host_frame == "shared":
        # N > 1, HOST result: the frame lives in POSIX shared memory, page-locked by every rank; each rank copies its
        # own bands device -> host into frame order over its own PCIe link (no gather on the device in this path)
        shared = pkg.tiles.SharedHostFrame(layout, dist, rank, count=2)
        shared.register(c)
        pipe.set_shared_host(shared, c)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pipe.step()
        pipe.drain()
        torch.cuda.synchronize()
        barrier()
        t_done = time.perf_counter()
        e2e_how = "every rank copies its bands D2H into a shared page-locked host frame (POSIX shm): all PCIe links in parallel, no device-side gather"
        checksum = int(shared.frame((args.steps - 1) % 2)[::64, ::64].astype(np.int64).sum()) if rank == 0 else 0
        pipe.set_shared_host(None, None)
        shared.close()
    else:
        pipe.host_frame = host_frame                  # rank 0 copies every gathered frame to pinned host memory
        for i in range(args.steps):
            pipe.step()
        pipe.drain()
        torch.cuda.synchronize()
        checksum = int(host_frame[::64, ::64].sum().item()) if rank == 0 else 0
    barrier()
    e2e_s = torch.tensor([(t_done if t_done is not None else time.perf_counter()) - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(e2e_s.item()) / args.steps
    e2e_value = rays / (e2e_ms / 1e3) / 1e6

    # what the end-to-end leg is bound by at N = 1: the same 4-byte-per-pixel frame copied device -> pinned host memory with
    # nothing else running (a diagnostic next to e2e, outside every timed region; failure only drops the field)
    d2h_alone = None
    if world == 1 and host_frame is not None:
        try:
            src = slab[:H]
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(3 + 10):
                if i == 3:
                    d0.record(stream)
                host_frame.copy_(src, non_blocking=True)
            d1.record(stream)
            torch.cuda.synchronize()
            d2h_ms = d0.elapsed_time(d1) / 10
            d2h_alone = {"ms_per_frame": d2h_ms, "gbs": W * H * 4 / (d2h_ms / 1e3) / 1e9}
        except Exception as e:                       # noqa: BLE001
            d2h_alone = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_how = measured_peak_gbs()
        ab = load_algorithmic_bytes(args.config, args.lights)
        key = "bytes_svo" if use_svo else "bytes_dense"
        algo_bytes = float(ab[key]) if ab else None
        st = c.stats()
        traffic = None
        tp = ROOT / "profiles" / "ncu_traffic.json"
        if world == 1 and tp.exists():
            t = json.loads(tp.read_text()).get(args.config, {}).get("vr_svo_kernel" if use_svo else "vr_dense_kernel")
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
        roofline = None
        if algo_bytes:
            achieved = algo_bytes / world / (kernel_ms / 1e3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_how, "kernel": "vr_svo_kernel" if use_svo else "vr_dense_kernel",
                        "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes / world,
                        "bytes_model": "P*(16+4) + 8*D_svo + 4*T from oracle counters (profiles/algorithmic_bytes_%s.json)" % args.config,
                        "node_bytes_fetched_per_launch": 16.0 * node_fetches / world}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_stride = args.cpu_row_stride * (8 if args.config == "c4" else 1)
            ref = CpuReference(scene, cpu_stride, args.lights)
            dt = min(ref.run(), ref.run()) if ref.lib is not None else ref.dt0
            if ref.lib is None and ref.kind == "port":
                dt = min(dt, ref.dt0)
            cpu = {"value": ref.rays / dt / 1e6, "unit": "Mrays/s", "cores": ref.threads, "kind": ref.kind,
                   "sample": f"every {cpu_stride}th row of the frame ({ref.rays} rays, {dt:.1f} s); {ref.note}"}
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": int(launches),
            "config": {"workload": f"{args.config}: {scene.n}^3 shell terrain SVO, {W}x{H}, primary + {args.lights} shadow light{'s (multi-light extension)' if args.lights > 1 else ''}, max_distance {scene.max_distance}" + (" (BASELINE configs[3], not the headline)" if args.config == "c4" else ""),
                       "mode": args.mode,
                       "walk": WALK_NOTES[args.walk] if use_svo else "dense DDA", "top_grid": "directed" if args.directed_grid else "undirected",
                       "other_walk_ms_per_frame": other_walk_ms, "sustained": sustained,
                       "kernel_variant": (f"persistent warps, refill_min {args.refill_min}, {args.ctas_per_sm} CTAs/SM" if args.persistent else "static 32x4 tiles, 128-thread CTAs, 8 CTAs/SM"), "parallelism": (f"tiles{world}: 2-D interleave of 32x4-pixel tiles ((tx + ty) % {world}), every rank's kernel stores its pixels in place into the root's frame over NVLink (CUDA IPC mapping), 1-element NCCL all_reduce as frame-complete signal, 3 frame buffers" if args.gather == "direct" else f"tiles{world}: interleaved {BAND_ROWS}-row bands, {'copy-engine push over NVLink (CUDA IPC) + 1-element NCCL all_reduce' if args.gather == 'p2p' else 'NCCL all_gather'} of frame k overlapped with rendering of frame k+1") if world > 1 else "1 GPU",
                       "l2": "per-frame streams (ray table 133 MB + image 33 MB) exceed the 126 MB L2; the octree stays L2-resident by design",
                       "rays_per_frame": rays, "primary_rays": primary, "shadow_rays": shadow, "primary_mpix_per_s": primary / (ms_per_step / 1e3) / 1e6,
                       "tree_nodes": int(st.native_nodes), "tree_bytes": int(st.native_bytes), "levels": int(st.levels),
                       "octree_broadcast_bytes": bcast_bytes, "scene_build_s": round(t_build, 2),
                       "octree_build": ({"where": "device (vr_build.cu) from the uploaded dense map", "ms": round(float(st.build_ms), 3),
                                         "map_read_ms": round(float(st.build_masks_ms), 3),
                                         "map_read_gbs": round(scene.n ** 3 / max(float(st.build_masks_ms), 1e-6) / 1e6, 1)}
                                        if st.build_masks_ms > 0 else
                                        {"where": "device (vr_build.cu) from the column tables: only the bricks that hold a voxel are materialised (no N^3 volume, no dense workspace)",
                                         "ms": round(float(st.build_ms), 3)} if st.build_ms > 0 else {"where": "host (column builder or broadcast)"}),
                       "dda_steps_per_frame": steps_total, "octree_lookups_per_frame": lookups, "frame_checksum": checksum, "device_frame_checksum": device_checksum},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 5 * 4 + 10 * 4 + 64 * 8,
                    "d2h_bytes_per_step": W * H * 4, "how": e2e_how, "d2h_copy_alone": d2h_alone},
            "clocks": clocks.summary(),
        }
        print(json.dumps(out))
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
