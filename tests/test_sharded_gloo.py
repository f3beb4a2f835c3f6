"""CPU, world_size 2, gloo: the host-side protocol of the sharded single-frame mode
(uw_slam_b200/sharded.py) -- contiguous partition, one all-reduce of the 32 sums per sweep,
identical termination and pose on every rank -- driven by a CPU stand-in for the device backend
built from the oracle's two GN halves.  The result must equal the unsharded oracle pose."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_range_without_overlap():
    from uw_slam_b200.sharded import partition
    for n in [0, 1, 7, 100, 790445]:
        for g in [1, 2, 3, 8]:
            parts = [partition(n, r, g) for r in range(g)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(g - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def _worker(rank, world, port, calib, seed, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import uw_oracle as O
    from uw_slam_b200 import synth
    from uw_slam_b200.sharded import estimate_pose_sharded, partition

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, seed)
    fp, fc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
    params = O.default_params(w, h, fx, fy, cx, cy)

    class OracleShard:
        """Same protocol as TrackerShardBackend, computed by the CPU oracle."""

        def begin(self, rank, nranks):
            self.rank, self.n = rank, nranks
            self.pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
            self.lvl, self.k, self.last = params.first_level, 0, 50000.0
            self.done = False

        def accumulate(self):
            lo, hi = partition(fp.cand[self.lvl].shape[0], self.rank, self.n)
            return torch.from_numpy(O.sweep_range(params, fp, fc, self.lvl, lo, hi, self.pose))

        def update(self, sums):
            brk, self.pose, self.last = O.gn_update(params, sums.numpy(), self.k, self.pose,
                                                    self.last)
            if brk:
                if self.lvl != 0:
                    self.pose = O.se3_scale_level(self.pose)
                self.lvl, self.k, self.last = self.lvl - 1, 0, 50000.0
                self.done = self.lvl < params.last_level
            else:
                self.k += 1
            return self.done

        def result(self):
            return self.pose, None

    pose, _, sweeps = estimate_pose_sharded(OracleShard())
    ref, _, tr = O.estimate_pose(params, fp, fc)
    q.put((rank, pose.tolist(), ref.tolist(), sweeps, len(tr)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("calib,seed", [("small", 0), ("tum", 1)])
def test_two_rank_sharded_loop_matches_unsharded_oracle(calib, seed):
    mp = pytest.importorskip("torch.multiprocessing")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, calib, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    (r0, pose0, ref0, sweeps0, ntr0), (r1, pose1, ref1, sweeps1, ntr1) = out
    assert pose0 == pose1                      # every rank holds the same pose, no broadcast
    assert pose0 == ref0                       # and it is the unsharded oracle's pose
    assert sweeps0 == sweeps1 == ntr0          # same number of sweeps as the reference loop
