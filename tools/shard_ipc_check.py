#!/usr/bin/env python
"""Two (or more) PROCESSES, one GPU each: the fused sharded Gauss-Newton kernel with its
in-kernel all-reduce over CUDA-IPC-mapped peer mailboxes (uwt_shard_ipc_export /
uwt_shard_ipc_connect), checked against the single-GPU kernel and the CPU oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/shard_ipc_check.py [--calib tum]

Rank 0 prints one JSON line.  Used by tests/test_gpu_multiproc.py."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from uw_slam_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calib", default="tum")
    ap.add_argument("--seeds", type=int, default=3)
    args = ap.parse_args()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import uw_slam_b200 as U
    from uw_slam_b200.sharded import (TrackerShardBackend, connect_fused, estimate_pose_sharded,
                                      estimate_pose_sharded_fused)
    rank, local_rank, world = (int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]),
                               int(os.environ["WORLD_SIZE"]))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h, fx, fy, cx, cy = synth.CALIB[args.calib]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2, device=local_rank)
    connect_fused(t)   # all-gather of the IPC handles, cudaIpcOpenMemHandle on every peer
    res = {"world": world, "calib": args.calib, "pairs": []}
    ok = True
    for seed in range(args.seeds):
        prev, cur, _, _ = synth.render_pair(args.calib, 40 + seed)   # numpy: same bytes everywhere
        t.AddFrames([0, 1], np.stack([prev, cur]))
        t.ApplyGradient([0])
        t.ObtainCandidatePoints([0])
        fpose, fst = estimate_pose_sharded_fused(t, 0, 1)
        npose, nst, sweeps = estimate_pose_sharded(TrackerShardBackend(t, 0, 1))
        single = t.EstimatePose([0], [1])[0]
        # every rank must hold the same pose: compare bit patterns across ranks
        mine = torch.from_numpy(fpose.view(np.int32).copy()).to("cuda")
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        same_ranks = all(bool(torch.equal(a, allp[0])) for a in allp)
        entry = {"seed": 40 + seed, "sweeps": sweeps,
                 "fused_equals_single_gpu": bool(np.array_equal(fpose, single)),
                 "nccl_equals_single_gpu": bool(np.array_equal(npose, single)),
                 "fused_same_on_all_ranks": same_ranks,
                 "fused_stats_equal_nccl_stats": list(fst.evaluations) == list(nst.evaluations)}
        if rank == 0:
            from oracle import uw_oracle as O
            p = O.default_params(w, h, fx, fy, cx, cy)
            op, _, _ = O.estimate_pose(p, O.FrameData(prev), O.FrameData(cur, with_candidates=False))
            entry["fused_equals_oracle"] = bool(np.array_equal(op, fpose))
        ok &= all(v for k, v in entry.items() if isinstance(v, bool))
        res["pairs"].append(entry)
        dist.barrier()
    res["ok"] = bool(ok)
    t.close()
    if rank == 0:
        print(json.dumps(res), file=real, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
