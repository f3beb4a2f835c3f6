"""The CUDA-IPC path of the fused sharded mode with REAL processes (one per GPU): skipped on a
box with fewer than two GPUs; `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multiproc`
runs it (log under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_fused_shard_over_cuda_ipc_processes(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
         "--master-port", str(29540 + world), os.path.join(ROOT, "tools", "shard_ipc_check.py"),
         "--calib", "tum", "--seeds", "2"],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    j = json.loads(lines[-1])
    assert j["world"] == world and j["ok"], j
    for e in j["pairs"]:
        assert e["fused_equals_single_gpu"] and e["fused_equals_oracle"] and \
            e["fused_same_on_all_ranks"] and e["nccl_equals_single_gpu"], e
