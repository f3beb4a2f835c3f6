"""A short run of the parity soak (tools/parity_soak.py): 256 tracks at 1280x1024 on the
dataflow kernel -- plain, occluded and sparse pairs -- every sweep's N_valid, sum r^2, error, A,
b, delta and pose against the oracle, bit for bit.  The long runs (100 032 + 20 032 tracks, 0
mismatches) are recorded under profiles/parity_soak*.json; this keeps the tool and the property
under test on every GPU run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_short_parity_soak(tmp_path):
    out = tmp_path / "soak.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_soak.py"),
                        "--tracks", "256", "--seed0", "7000000", "--out", str(out)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    j = json.loads(out.read_text())
    assert j["tracks"] == 256 and j["sweeps"] > 2000
    assert j["sweep_mismatches"] == 0 and j["pose_mismatches"] == 0, j["examples"]
    assert j["trace_length_mismatches"] == 0
    assert all(v > 0 for v in j["variants"].values())


def test_committed_soak_records_are_clean():
    """CPU: the committed long-run records say what the docs say."""
    for name, tracks in (("parity_soak.json", 100032), ("parity_soak_r02_final.json", 20032),
                         # the end-of-round kernels (16384-record tasks, late-sweep rule, rewritten
                         # scatter): every sweep at 64 problems per call, final poses at 256
                         ("parity_soak_r02s_traced.json", 8192),
                         ("parity_soak_r02s_batch256.json", 16384),
                         ("parity_soak_r02v_50k.json", 50048)):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            j = json.load(f)
        assert j["tracks"] == tracks and j["sweeps"] > 10 * tracks
        assert j["sweep_mismatches"] == 0 and j["pose_mismatches"] == 0
        assert sum(j["field_mismatches"].values()) == 0
