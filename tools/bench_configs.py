#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 1, 3, 4); bench.py is config 2, the headline.

  --workload euroc_seq   config 1: 752x480 EuRoC-calibration sequence, ONE stream, frame-to-
                         keyframe (re-keyed every 8 frames): single-problem latency
  --workload batch8192   config 3: 8192 independent 640x480 pairs, block-partitioned over the
                         ranks (strong scaling, no comms), processed in resident chunks
  --workload hypotheses  config 3, hypothesis variant: one 640x480 pair, 8192 initial poses
  --workload shard4k     config 4: one 3840x2160 pair, GN loop sharded over the ranks with one
                         NCCL all-reduce of the 32 normal-equation sums per sweep

Launch like bench.py (python tools/bench_configs.py ... or torchrun for N > 1).  One JSON line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

# keep stdout for the one JSON line: NCCL banners / debug output go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from uw_slam_b200 import synth  # noqa: E402


class Args:
    """Defaults of the command line below, for callers that import the workload functions
    (bench.py's `secondary` block)."""

    def __init__(self, **kw):
        self.pairs, self.chunk, self.reps, self.check = 8192, 1024, 20, False
        self.flags, self.weights, self.depth = 0, 0, 0
        self.__dict__.update(kw)


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def make_tracker(calib, device, **cfg):
    import uw_slam_b200 as U
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        device=device, **cfg)
    return t


def euroc_seq(args, torch, dist, rank, local_rank, world):
    if rank != 0:   # a single camera stream does not shard: the other ranks have nothing to do
        return None
    calib, n_frames, rekey = "euroc", 64, 8
    w, h = synth.CALIB[calib][:2]
    frames, _, _ = synth.render_sequence(calib, 11, n_frames, rot=1.5e-3, trans=1.5e-3)
    host = torch.from_numpy(np.stack(frames)).pin_memory()
    t = make_tracker(calib, local_rank, max_frames=2)
    KEY, CUR = 0, 1

    def run(collect):
        poses, sweeps = [], 0
        t.AddFramesHostPtr([KEY], host[0].data_ptr(), w, w * h)
        t.ApplyGradient([KEY])
        t.ObtainCandidatePoints([KEY])
        for i in range(1, n_frames):
            t.AddFramesHostPtr([CUR], host[i].data_ptr(), w, w * h)
            pose, st = t.EstimatePose([KEY], [CUR], return_stats=True)   # keyframe -> frame i
            if collect:
                poses.append(pose[0])
                sweeps += sum(st[0].evaluations)
            if i % rekey == 0:  # the current frame becomes the keyframe
                t.AddFramesHostPtr([KEY], host[i].data_ptr(), w, w * h)
                t.ApplyGradient([KEY])
                t.ObtainCandidatePoints([KEY])
        return poses, sweeps

    run(False)
    t.synchronize()
    t.profile(True)
    t0 = time.perf_counter()
    poses, sweeps = run(True)
    t.synchronize()
    dt = time.perf_counter() - t0
    prof = t.profile_read()
    # CPU oracle on the same sequence (bounded: it is only 63 tracks)
    from oracle import uw_oracle as O
    p = O.default_params(*synth.CALIB[calib], accum_mode=0)
    t1 = time.perf_counter()
    key = O.FrameData(frames[0])
    same = True
    for i in range(1, n_frames):
        cur = O.FrameData(frames[i], with_candidates=False)
        op, _, _ = O.estimate_pose(p, key, cur)
        same &= bool(np.array_equal(op, poses[i - 1]))
        if i % rekey == 0:
            key = O.FrameData(frames[i])
    cpu_dt = time.perf_counter() - t1
    est_ms, est_l = prof["estimate"]
    t.close()
    return {"metric": "pose-tracks/sec, single stream 752x480 (frame-to-keyframe)",
            "value": (n_frames - 1) / dt, "unit": "tracks/s", "tracks": n_frames - 1,
            "ms_per_track_e2e": 1e3 * dt / (n_frames - 1),
            "estimate_kernel_us_per_track": 1e3 * est_ms / max(est_l, 1),
            "us_per_gn_sweep": 1e3 * est_ms / max(sweeps, 1), "sweeps_per_track":
            sweeps / (n_frames - 1),
            "cpu_oracle_tracks_per_s": (n_frames - 1) / cpu_dt,
            "poses_bit_identical_to_oracle": same, "n_gpus": 1}


def batch8192(args, torch, dist, rank, local_rank, world):
    calib, total = "tum", args.pairs
    w, h = synth.CALIB[calib][:2]
    dev = torch.device("cuda", local_rank)
    lo, hi = total * rank // world, total * (rank + 1) // world   # pair i -> GPU floor(i*G/B)
    mine = hi - lo
    chunk = min(args.chunk, mine)
    t = make_tracker(calib, local_rank, max_frames=2 * chunk, flags=args.flags,
                     weight_mode=args.weights, depth_mode=args.depth)
    depth = None
    if args.depth:
        # one smooth synthetic depth map with holes, shared by all pairs (resident in HBM)
        ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev),
                                indexing="ij")
        dm = (9000 + 4000 * torch.sin(xs * 0.013) * torch.cos(ys * 0.017)).to(torch.int32)
        dm[(xs * 7 + ys * 13) % 17 == 0] = 0
        depth = dm.to(torch.uint16).contiguous()
    # resident inputs: all of this rank's pairs, rendered on the device
    prev = torch.empty((mine, h, w), dtype=torch.uint8, device=dev)
    cur = torch.empty_like(prev)
    for a in range(0, mine, 256):
        b = min(a + 256, mine)
        p_, c_ = synth.render_pairs_torch(calib, list(range(lo + a, lo + b)), dev)
        prev[a:b], cur[a:b] = p_, c_
    ps, cs = list(range(chunk)), list(range(chunk, 2 * chunk))
    out = np.empty((mine, 7), np.float32)
    torch.cuda.synchronize()   # torch's streams are not ordered with the library's stream

    def run():
        for a in range(0, mine, chunk):
            n = min(chunk, mine - a)
            t.AddFramesDevice(ps[:n], prev[a].data_ptr())
            t.AddFramesDevice(cs[:n], cur[a].data_ptr())
            if depth is not None:
                # frame_stride 0: every slot gets the same (device-resident) depth frame
                t._check(t._lib.uwt_upload_depth_frames(t._h, n, (C.c_int * n)(*ps[:n]),
                                                        depth.data_ptr(), w * 2, 0))
            t.ApplyGradient(ps[:n])
            t.ObtainCandidatePoints(ps[:n])
            out[a:a + n] = t.EstimatePose(ps[:n], cs[:n])

    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run()
    t.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    # device time per kernel class of one more pass (events; not part of the timed figure)
    t.profile(True)
    run()
    prof = {k: v[0] for k, v in t.profile_read().items()}
    t.profile(False)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res = {"metric": "pose-tracks/sec, 8192 independent 640x480 pairs (pyramids x2, gradient, "
                     "candidates, estimate per pair; inputs resident in HBM)",
           "value": total / float(dt.item()), "unit": "tracks/s", "pairs": total,
           "n_gpus": world, "scaling": "strong", "chunk": chunk, "flags": args.flags,
           "weights": ["identity", "tukey_mad", "huber"][args.weights], "depth_mode": args.depth,
           "kernel_ms_per_pass_rank0": prof}
    if rank == 0:
        # spot-check 4 pairs against the oracle
        from oracle import uw_oracle as O
        p = O.default_params(*synth.CALIB[calib], weight_mode=args.weights)
        same = True
        dnp = depth.cpu().numpy() if depth is not None else None
        for i in [0, 1, mine // 2, mine - 1]:
            a, b = prev[i].cpu().numpy(), cur[i].cpu().numpy()
            op, _, _ = O.estimate_pose(p, O.FrameData(a, depth=dnp, depth_mode=args.depth),
                                       O.FrameData(b, with_candidates=False))
            same &= bool(np.array_equal(op, out[i]))
        res["spot_check_bit_identical_to_oracle"] = same
    t.close()
    return res


def hypotheses(args, torch, dist, rank, local_rank, world):
    """Config 3, hypothesis variant: ONE 640x480 pair, `--pairs` initial-pose hypotheses split
    over the ranks (strong scaling, no comms).  Every hypothesis is an independent problem that
    shares the frame slots (prev = slot 0, cur = slot 1)."""
    from oracle import uw_oracle as O
    total, calib = args.pairs, "tum"
    lo, hi = total * rank // world, total * (rank + 1) // world
    mine = hi - lo
    chunk = min(args.chunk, mine)
    t = make_tracker(calib, local_rank, max_frames=max(2, chunk), flags=args.flags)
    prev, cur, _, _ = synth.render_pair(calib, 5)
    t.AddFrames([0, 1], np.stack([prev, cur]))
    t.ApplyGradient([0])
    t.ObtainCandidatePoints([0])
    rng = np.random.default_rng(1234)
    tang = (rng.normal(size=(total, 6)) * 2e-3).astype(np.float32)
    init = np.stack([O.se3_exp(a) for a in tang[lo:hi]])
    out = np.empty((mine, 7), np.float32)

    def run():
        for a in range(0, mine, chunk):
            n = min(chunk, mine - a)
            out[a:a + n] = t.EstimatePose([0] * n, [1] * n, init_poses=init[a:a + n])

    run()
    t.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run()
    t.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64,
                      device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res = {"metric": "pose estimates/sec, one 640x480 pair, %d initial-pose hypotheses" % total,
           "value": total / float(dt.item()), "unit": "estimates/s", "hypotheses": total,
           "n_gpus": world, "scaling": "strong", "chunk": chunk}
    if rank == 0:
        p = O.default_params(*synth.CALIB[calib])
        rp, rc = O.FrameData(prev), O.FrameData(cur, with_candidates=False)
        res["spot_check_bit_identical_to_oracle"] = all(
            bool(np.array_equal(O.estimate_pose(p, rp, rc, init_pose=init[i])[0], out[i]))
            for i in (0, mine // 2, mine - 1))
    return res


def shard4k(args, torch, dist, rank, local_rank, world):
    """Config 4: ONE 3840x2160 pair, the candidate list of every level split over the ranks.
    Three forms on the same pair, every one checked bit for bit against the others:
      fused   : one persistent kernel per rank, all-reduce of the 32 fp64 sums through peer
                (CUDA-IPC mapped) mailboxes inside the kernel
      nccl    : per sweep accumulate kernel -> NCCL all-reduce -> update kernel, host loop
      1 GPU   : the SAME fused kernel as a single-rank instance on rank 0's GPU alone
                (uwt_estimate_pose's path for one large frame)"""
    from uw_slam_b200.sharded import (TrackerShardBackend, connect_fused, estimate_pose_sharded,
                                      estimate_pose_sharded_fused)
    calib = "uhd"
    w, h = synth.CALIB[calib][:2]
    dev = torch.device("cuda", local_rank)
    # every rank must hold the same bytes: rank 0 renders, the others receive
    pair = torch.empty((2, h, w), dtype=torch.uint8, device=dev)
    if rank == 0:
        p_, c_ = synth.render_pairs_torch(calib, [21], dev)
        pair[0], pair[1] = p_[0], c_[0]
    if world > 1:
        dist.broadcast(pair, src=0)
    torch.cuda.synchronize()   # torch's streams are not ordered with the library's stream
    t = make_tracker(calib, local_rank, max_frames=2)
    t.AddFramesDevice([0, 1], pair.data_ptr())
    t.ApplyGradient([0])
    t.ObtainCandidatePoints([0])
    reps = args.reps

    def timed(fn, sync):
        fn()
        sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        r2 = sync()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()), (r2 if r2 is not None else r)

    # NCCL + host loop
    backend = TrackerShardBackend(t, 0, 1)
    dt_nccl, (pose, stats, sweeps) = timed(lambda: estimate_pose_sharded(backend),
                                           lambda: t.synchronize())
    # fused in-kernel all-reduce over the ranks
    connect_fused(t)
    dt_fused, (fpose, _) = timed(lambda: t.ShardEstimateFusedAsync(0, 1),
                                 lambda: t.ShardEstimateFusedWait())
    # the same kernel on one GPU (every rank runs its own copy; rank 0's figure is reported)
    t.EstimatePose([0], [1])
    t.synchronize()
    t1 = time.perf_counter()
    for _ in range(reps):
        ref = t.EstimatePose([0], [1])
    dt_single = (time.perf_counter() - t1) / reps
    # device time of the kernels alone (CUDA events around the launches): what sharding changes
    def kernel_ms(fn):
        t.synchronize()
        if world > 1:
            dist.barrier()
        t.profile(True)
        for _ in range(10):
            fn()
        ms = t.profile_read()["estimate"][0] / 10
        t.profile(False)
        return ms
    k_fused = kernel_ms(lambda: (t.ShardEstimateFusedAsync(0, 1), t.ShardEstimateFusedWait()))
    k_single = kernel_ms(lambda: t.EstimatePose([0], [1]))
    res = {"metric": "GN pose estimate of one 3840x2160 pair, candidate list sharded over ranks",
           "kernel_ms_fused_peer_allreduce": k_fused,
           "kernel_ms_same_fused_kernel_1gpu": k_single,
           "kernel_speedup_fused_vs_1gpu": k_single / k_fused if k_fused else None,
           "n_gpus": world, "sweeps": sweeps, "points_per_level": list(stats.n_points)[:5],
           "ms_per_estimate_fused_peer_allreduce": 1e3 * dt_fused,
           "us_per_sweep_fused": 1e6 * dt_fused / sweeps,
           "ms_per_estimate_nccl_host_loop": 1e3 * dt_nccl,
           "us_per_sweep_nccl_host_loop": 1e6 * dt_nccl / sweeps,
           "ms_per_estimate_same_fused_kernel_1gpu": 1e3 * dt_single,
           "us_per_sweep_same_fused_kernel_1gpu": 1e6 * dt_single / sweeps,
           "speedup_fused_vs_same_kernel_1gpu": dt_single / dt_fused,
           "fused_pose_equals_1gpu_kernel": bool(np.array_equal(fpose, ref[0])),
           "nccl_pose_equals_1gpu_kernel": bool(np.array_equal(pose, ref[0])),
           "collective": ("none (1 rank)" if world == 1 else
                          "fused: peer-mailbox all-reduce of 32 fp64 inside the kernel; "
                          "nccl: NCCL all-reduce of 32 fp64 per sweep")}
    if rank == 0 and args.check:
        from oracle import uw_oracle as O
        prev, cur = pair[0].cpu().numpy(), pair[1].cpu().numpy()
        p = O.default_params(*synth.CALIB[calib])
        op, _, _ = O.estimate_pose(p, O.FrameData(prev), O.FrameData(cur, with_candidates=False))
        res["pose_bit_identical_to_oracle"] = bool(np.array_equal(op, fpose))
    t.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["euroc_seq", "batch8192", "hypotheses", "shard4k"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--pairs", type=int, default=8192)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="uwt_config.flags (4 = cluster kernel)")
    ap.add_argument("--weights", type=int, default=0, help="0 identity, 1 Tukey/MAD, 2 Huber")
    ap.add_argument("--depth", type=int, default=0,
                    help="depth input: 0 mono, 1 reference (at<uchar>), 2 at<ushort>")
    args = ap.parse_args()
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")   # banners (NCCL ...) go to stderr, the JSON here
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = {"euroc_seq": euroc_seq, "batch8192": batch8192, "hypotheses": hypotheses,
           "shard4k": shard4k}[args.workload](
        args, torch, dist, rank, local_rank, world)
    res["workload"] = args.workload
    if rank == 0:
        print(json.dumps(res), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
