"""The reference arm of bench.py (CPU port on the host cores) prints ONE JSON line with the keys the driver reads; run
here with the smallest sample (1 warm-up + 1 timed frame of the c2 scene)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun would set: the port must still use every host thread
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("frames/sec") and "c2" in d["config"]["workload"] and d["dtype"] == "f32"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "data"):
        assert k in d, k
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["sample"]
    n_host = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == n_host
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_describe_the_same_workload():
    """`config` is built by one function for both arms (the driver's same_config check), and every rank times the same set
    of frames at every world size (per-rank work independent of N)."""
    sys.path.insert(0, ROOT)
    import bench

    cfg = bench.base_config()
    assert cfg["workload"].startswith("c2") and cfg["gaussians"] == 1_000_000 and (cfg["width"], cfg["height"]) == (1920, 1080)
    base = sorted(bench.bench_frames(0, 25))
    for rank in range(8):
        fr = bench.bench_frames(rank, 25)
        assert sorted(fr) == base and len(set(fr)) == 25
    assert bench.bench_frames(0, 25) != bench.bench_frames(1, 25)  # rotated, not in lock step
