"""bench.py contract (CPU part): the reference arm prints exactly one JSON line on stdout with
the keys the driver reads, whatever libraries print banners; rank != 0 prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "4",
             "--cpu-sequences-1t", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "tracks/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("pose-tracks/sec @1280x1024")
    assert j["steps"] == 1 and j["warmup"] == 0 and j["n_gpus"] == 1 and j["value"] > 0
    assert j["config"]["workload"] == "tum_mono_1280x1024_seq"
    assert j["config"]["sequences_per_gpu"] == 4 and j["config"]["tracks_per_step"] == 4
    # the config object is built by the same function for both arms
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.workload_config(1, 4, 1, 0)
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert (e["value"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]) == (j["value"], 0, 0)


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
            {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
