"""CPU tests of bench.py's host-side logic (no GPU): the reference arm's JSON line and its behaviour under a torchrun-like
environment (the launcher exports OMP_NUM_THREADS=1 and starts one process per rank), and the clock sampler's contract that
nothing slow happens between the barrier in front of a timed region and the first launch."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def _reference_arm(extra_env: dict, *flags: str) -> subprocess.CompletedProcess:
    env = dict(os.environ)
    env.update(extra_env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "3", *flags],
                          env=env, capture_output=True, text=True, timeout=300, cwd=str(ROOT))


def test_reference_arm_line_under_a_launcher_that_exports_one_omp_thread():
    """torch.distributed.run exports OMP_NUM_THREADS=1: the reference arm must still use every host core and print the team
    size it actually ran with (the arm ran single-threaded while printing "cores": 32)."""
    r = _reference_arm({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout carries exactly one JSON line"
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == bench.METRIC and j["unit"] == "Mrays/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 2 and j["steps"] == 2 and j["warmup"] == 3 and j["gpu_launches"] == 0
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["value"] == j["value"] and cb["kind"] in ("reference", "port") and "row" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)), "the team size OpenMP used, not the launcher's OMP_NUM_THREADS=1"


def test_reference_arm_other_ranks_exit_quietly():
    r = _reference_arm({"OMP_NUM_THREADS": "1", "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_clock_sampler_is_parked_by_prepare_and_released_by_entering():
    """prepare() does everything slow (NVML, thread start-up); entering the context only sets an event; nothing is sampled
    before that; leaving joins the thread -- also when the context was never entered."""
    s = bench.ClockSampler(0, enabled=True, delay_first=True).prepare()
    assert s._t is not None and s._t.is_alive() and not s._go.is_set()
    time.sleep(0.02)
    assert s.samples == [], "parked: no query before the timed region starts"
    t0 = time.perf_counter()
    with s as clocks:
        dt_enter = time.perf_counter() - t0
        assert s._go.is_set()
    assert dt_enter < 5e-3, f"entering the sampler took {dt_enter * 1e3:.2f} ms"
    assert not s._t.is_alive()
    assert set(clocks.summary()) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}

    never_entered = bench.ClockSampler(0).prepare()
    never_entered.__exit__(None, None, None)
    assert not never_entered._t.is_alive()

    off = bench.ClockSampler(0, enabled=False).prepare()
    with off as clocks:
        pass
    assert off._t is None and clocks.summary()["samples"] == 0


def test_bench_defaults_are_the_contract():
    """no flags: one GPU, a K / W that finish within minutes, W >= 3"""
    src = (ROOT / "bench.py").read_text()
    assert 'add_argument("--gpus", type=int, default=1)' in src
    assert "args.warmup = max(args.warmup, 3)" in src
    assert bench.METRIC.startswith("Mrays/s")


REQUIRED_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                 "data", "gpu_launches", "config", "roofline", "cpu_baseline", "e2e", "clocks"}


def _mock_bench(cmd_prefix, *flags, timeout=600):
    r = subprocess.run([*cmd_prefix, str(ROOT / "tests" / "bench_mock_device.py"), "--config", "c1", "--steps", "6", "--warmup", "3", *flags],
                       capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, f"exactly one JSON line on stdout, got {len(lines)}"
    return json.loads(lines[0])


def test_bench_control_flow_one_gpu_on_a_fake_device():
    """bench.py's N = 1 path, start to JSON line, against tests/bench_mock_device.py (no rendering, virtual event clock): every
    key of the contract is there, the step count is the launch count, the roofline and e2e objects are well-formed."""
    j = _mock_bench([sys.executable])
    assert REQUIRED_KEYS <= set(j), REQUIRED_KEYS - set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 6 and j["gpu_launches"] == 6 and j["ms_per_step"] == 0.5
    assert j["config"]["workload"].startswith("c1:") and j["config"]["sustained"]["frames"] >= 6
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(j["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "d2h_copy_alone"} <= set(j["e2e"])
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1


def test_bench_control_flow_two_ranks_on_a_fake_device():
    """the N > 1 path (run_mgpu) under torch.distributed.run with two gloo ranks: the collectives match on both ranks, rank 0
    alone prints the line, and it carries the per-rank times, the start skew after the alignment and the sustained leg."""
    port = 29000 + os.getpid() % 2000
    j = _mock_bench([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                     "--master-port", str(port)], "--gpus", "2")
    assert REQUIRED_KEYS <= set(j), REQUIRED_KEYS - set(j)
    assert j["n_gpus"] == 2 and j["steps"] == 6 and j["gpu_launches"] == 6
    cfg = j["config"]
    assert len(cfg["per_rank_ms_per_frame"]) == 2 and cfg["start_skew_us"] >= 0.0
    assert cfg["sustained"]["frames"] >= 6 and cfg["sustained"]["ms_per_frame"] > 0
    assert j["e2e"]["value"] > 0 and j["cpu_baseline"] is None
