#!/usr/bin/env python
"""bench.py -- pose-tracks/sec of the direct photometric tracker @1280x1024 (BASELINE.json).

Workload (config.workload = "tum_mono_1280x1024_seq"): B independent camera sequences at
1280x1024 (TUM-mono pinhole), 5-level pyramid, levels 4..1 optimised (the reference's
"4-level" tracking), all high-gradient pixels as candidates.  One STEP advances every
sequence by one frame = B pose-tracks; one track is exactly the reference's per-frame work
in the direct order:
    pyramid(cur) -> EstimatePose(prev, cur) -> ApplyGradient(cur) -> ObtainCandidatePoints(cur)
`value`  : tracks/s with the new frames already resident in HBM when the step starts.
`e2e`    : the same loop through the C ABI with frames in pinned HOST memory: the H2D copy of
           every frame and the D2H read of every pose are inside the timed region.
`--impl reference`: the CPU oracle (restatement of the reference, oracle/) on all host cores, on
           the SAME sequences (seeds, frames, K, W) as one GPU of `--impl ours`: identical
           `config`.  Every rank of `--impl ours` tracks the same B seeded sequences (weak
           scaling with fixed per-GPU work); `per_rank` lists every rank's own time, kernel
           table, sweeps and clocks, so a slow rank is attributable.
`secondary`: BASELINE configs 1 / 3 / 4 (EuRoC single stream, 8192 pairs split over the ranks,
           one 3840x2160 pair sharded over the ranks) measured after the timed loops.

Launch: python bench.py [--gpus N --steps K --warmup W]   (N > 1: under torchrun, one rank
per GPU, weak scaling: every rank tracks its own B sequences, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# keep stdout for the one JSON line: NCCL banners / debug output go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from uw_slam_b200 import synth  # noqa: E402

CALIB = "tum_mono"
METRIC = "pose-tracks/sec @1280x1024 4-level pyramid"
ROT = TRANS = 5e-3  # per-frame motion magnitude
BYTES_PER_POINT = 10  # (x,y) 4 B + I1 1 B + gx,gy 4 B + I2 gather 1 B


def claim_stdout():
    """NCCL and friends print banners on fd 1; the contract is ONE JSON line on stdout.  Point
    fd 1 at stderr for the whole run and return a file on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned host buffer is touched,
    so the frames every step uploads live in memory local to that GPU's PCIe root (first-touch).
    With 8 ranks on a two-socket host the H2D streams otherwise cross the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml pads the domain to 8 hex digits
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def gen_sequences(torch, device, seeds, n_frames, chunk=16):
    """u8 tensor [n_frames, B, H, W]: frame i of sequence b is texture(seed b) seen after i
    steps of that sequence's seeded motion."""
    w, h = synth.CALIB[CALIB][:2]
    calib = synth.CALIB[CALIB]
    B = len(seeds)
    out = torch.empty((n_frames, B, h, w), dtype=torch.uint8, device=device)
    ys, xs = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float32),
                            torch.arange(w, device=device, dtype=torch.float32), indexing="ij")
    for b0 in range(0, B, chunk):
        bs = seeds[b0:b0 + chunk]
        tps = [synth.texture_params(s) for s in bs]
        kx, ky, amp, ph = [torch.tensor(np.stack([tp[j] for tp in tps]), device=device,
                                        dtype=torch.float32)[:, :, None, None] for j in range(4)]
        for i in range(n_frames):
            Hs = []
            for s in bs:
                om, t = synth.motion(s, ROT, TRANS)
                Hs.append(np.eye(3) if i == 0 else synth.homography_inv(calib, om * i, t * i))
            Hi = torch.tensor(np.stack(Hs), device=device, dtype=torch.float32)
            d = Hi[:, 2, 0, None, None] * xs + Hi[:, 2, 1, None, None] * ys + Hi[:, 2, 2, None, None]
            qx = (Hi[:, 0, 0, None, None] * xs + Hi[:, 0, 1, None, None] * ys +
                  Hi[:, 0, 2, None, None]) / d
            qy = (Hi[:, 1, 0, None, None] * xs + Hi[:, 1, 1, None, None] * ys +
                  Hi[:, 1, 2, None, None]) / d
            v = torch.zeros_like(qx)
            for k in range(synth.N_WAVES):
                v += amp[:, k] * torch.sin(kx[:, k] * qx + ky[:, k] * qy + ph[:, k])
            out[i, b0:b0 + len(bs)] = torch.clamp(torch.round(127.5 + v), 0, 255).to(torch.uint8)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(batch):
    """DRAM bytes per launch of the estimate kernel from a committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            j = json.load(f)
        e = j.get("estimate_kernel", {})
        if int(e.get("batch", -1)) == int(batch):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def workload_config(world, B, K, W):
    """The workload both arms run, as the `config` object of the JSON line (identical keys and
    values in `--impl ours` and `--impl reference` for the same --gpus/--steps/--warmup/--batch)."""
    w, h = synth.CALIB[CALIB][:2]
    return {"workload": "tum_mono_1280x1024_seq", "sequences_per_gpu": B,
            "tracks_per_step": world * B, "tracked_frames_per_sequence": W + K,
            "seeds": "sequence i of every GPU = texture/motion seed i (frame 0 prepared "
                     "before the timed span, then one track per step)",
            "levels": 5, "optimised_levels": "4..1",
            "candidates": "all pixels with g > mean+20", "weights": "identity",
            "l2_policy": "inputs larger than L2 (%.0f MB of new frames + %.1f GB working set "
                         "per step and GPU)" % (B * w * h / 1e6, 2 * B * 21e6 / 1e9),
            "parallelism": "independent sequences, %d per GPU, no comms" % B}


class CpuTrackers:
    """The oracle's reference-shaped loop (oracle/uw_oracle.cpp uwo_stream_*: the previous
    frame is kept between calls like System keeps it, main_uw_slam.cpp:139-151) for many
    sequences on a pool of host threads, advanced one frame per step.  Frame 0 of every
    sequence is prepared in the constructor, outside any timed span."""

    def __init__(self, frames0, threads, weight_mode=0):
        from concurrent.futures import ThreadPoolExecutor
        from oracle import uw_oracle as O
        O.build()
        O.lib()
        w, h, fx, fy, cx, cy = synth.CALIB[CALIB]
        self.n = frames0.shape[0]
        self.threads = max(1, min(threads, self.n))
        self.params = [O.default_params(w, h, fx, fy, cx, cy, threads=1, weight_mode=weight_mode)
                       for _ in range(self.n)]
        self.pool = ThreadPoolExecutor(max_workers=self.threads)
        self.groups = [list(range(i, self.n, self.threads)) for i in range(self.threads)]
        self.streams = [None] * self.n

        def mk(ids):
            for s in ids:
                self.streams[s] = O.Stream(self.params[s], frames0[s])
        list(self.pool.map(mk, self.groups))

    def step(self, frames):
        """frames: u8 [n, H, W] (the next frame of every sequence).  Returns (poses [n,7],
        wall seconds of the step).  ctypes releases the GIL: the threads run in parallel."""
        poses = np.empty((self.n, 7), np.float32)

        def work(ids):
            for s in ids:
                poses[s] = self.streams[s].track(frames[s])
        t0 = time.perf_counter()
        list(self.pool.map(work, self.groups))
        return poses, time.perf_counter() - t0

    def close(self):
        for s in self.streams:
            if s is not None:
                s.close()
        self.pool.shutdown()


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's tracker (oracle/; the
    reference itself is not buildable here) on the SAME sequences as rank 0 of
    `--impl ours` (same seeds, same frames, same K and W), one step = every sequence advanced
    by one frame, all host threads busy, one single-threaded tracker per sequence."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    B, K, W = args.batch, args.steps, args.warmup
    seeds = list(range(B))
    fr = gen_sequences(torch, dev, seeds, 1 + W + K).cpu().numpy()      # [frames, B, H, W]
    # the single-thread figure first (the reference is single-threaded): a bounded sample of the
    # same sequences, same steps, on an otherwise idle host
    n1 = min(B, args.cpu_sequences_1t)
    trk1 = CpuTrackers(fr[0][:n1], 1)
    t1 = 0.0
    for i in range(1, 1 + W + K):
        _, dt = trk1.step(fr[i][:n1])
        if i > W:
            t1 += dt
    trk1.close()
    trk = CpuTrackers(fr[0], threads)
    times = []
    for i in range(1, 1 + W + K):
        _, dt = trk.step(fr[i])
        if i > W:
            times.append(dt)
    trk.close()
    total = sum(times)
    value = B * K / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * total / max(K, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64acc", "data": "synthetic",
        "config": workload_config(args.gpus, B, K, W),
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": trk.threads, "kind": "port",
                         "sample": "the %d sequences of one GPU x %d steps (%d tracks, %.1f s): "
                                   "oracle port of the reference tracker, one single-threaded "
                                   "tracker per sequence on %d host threads; at --gpus N > 1 the "
                                   "job has N x %d sequences per step, of which this is one "
                                   "GPU's share" % (B, K, B * K, total, trk.threads, B),
                         "single_thread_value": n1 * K / t1 if t1 > 0 else None,
                         "single_thread_sample": "%d of those sequences x %d steps" % (n1, K)},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def h2d_ceiling(torch, dist, dev, host, world, reps=8):
    """What the host can feed: every rank copies `reps` x one step's frames (the same pinned
    buffers the e2e loop uploads) to its GPU with plain cudaMemcpyAsync, all ranks at once.
    Returns GB/s of this rank."""
    dst = torch.empty_like(host[0], device=dev)
    dst.copy_(host[0], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dst.copy_(host[1 + i % (host.shape[0] - 1)], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * host[0].numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def run_ours(args):
    import torch
    import torch.distributed as dist
    import uw_slam_b200 as U
    from uw_slam_b200 import _lib as L

    rank, local_rank, world = dist_env()
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w, h, fx, fy, cx, cy = synth.CALIB[CALIB]
    B, K, W = args.batch, args.steps, args.warmup
    n_frames = 1 + W + K + 1  # +1: the e2e loop uploads one frame ahead

    # ---- synthetic sequences (device + pinned host copies) ----
    # Weak scaling with fixed per-GPU work: every rank tracks the same B seeded sequences, so
    # the per-rank times differ only by what the hardware does (and --impl reference tracks
    # exactly one GPU's share of the job).
    seeds = list(range(B))
    frames = gen_sequences(torch, dev, seeds, n_frames)          # [n_frames, B, H, W]
    host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(frames)
    torch.cuda.synchronize()

    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * B, device=local_rank, cluster_size=args.cluster,
                        flags=(L.FLAG_DMMA_ACCUM if args.dmma_accum else 0) |
                        (L.FLAG_CLUSTER_KERNEL if args.cluster_kernel else 0) |
                        (L.FLAG_LAZY_LEVELS if args.lazy_levels else 0),
                        weight_mode=args.weights)
    stream = torch.cuda.ExternalStream(t.stream_ptr(), device=dev)
    slots_a, slots_b = list(range(B)), list(range(B, 2 * B))
    frame_bytes = w * h

    def prime():
        t.AddFramesDevice(slots_a, frames[0].data_ptr())
        t.ApplyGradient(slots_a)
        t.ObtainCandidatePoints(slots_a)

    def upload(i, cur, from_host):
        if from_host:
            t.AddFramesHostPtr(cur, host[i].data_ptr(), w, frame_bytes)
        else:
            t.AddFramesDevice(cur, frames[i].data_ptr())

    def track(prev, cur):
        t.EstimatePoseAsync(prev, cur)
        t.ApplyGradient(cur)
        t.ObtainCandidatePoints(cur)

    def barrier():
        t.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fetch():
        st = (L.TrackStats * B)()
        out = np.empty((B, 7), np.float32)
        t._check(t._lib.uwt_fetch_poses(t._h, B, out.ctypes.data_as(L._fp), st))
        return out, st

    def timed_loop(from_host, fetch_poses):
        """Software-pipelined: the frames of step i+1 are handed to the library (H2D on its
        copy stream when they come from the host) while step i is still being tracked, and
        the poses of step i are read after that.  Every timed step contains exactly one
        upload, one track and (e2e) one pose read."""
        prime()
        prev, cur = slots_a, slots_b
        stats_acc, all_poses = [], []
        upload(1, cur, from_host)
        for i in range(1, 1 + W):
            track(prev, cur)
            upload(i + 1, prev, from_host)   # next step's frames go where `prev` lives
            if fetch_poses:
                all_poses.append(fetch()[0])
            prev, cur = cur, prev
        barrier()
        t.profile(True)
        l0 = t.launch_count(include_aux=True)
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        e0.record(stream)
        for i in range(1 + W, 1 + W + K):
            track(prev, cur)
            upload(i + 1, prev, from_host)
            if fetch_poses:
                stats_acc.append(fetch())
                all_poses.append(stats_acc[-1][0])
            prev, cur = cur, prev
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop()
        ms_own = e0.elapsed_time(e1)
        prof = t.profile_read()
        launches = t.launch_count(include_aux=True) - l0
        t.profile(False)
        ms = ms_own
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, ms_own, wall, prof, launches, clocks, stats_acc, all_poses

    # ---- value: inputs resident in HBM ----
    ms, ms_own, wall, prof, launches, clocks, _, _ = timed_loop(from_host=False, fetch_poses=False)
    # ---- e2e: pinned host frames in, poses out, every step ----
    ms_e, ms_e_own, wall_e, prof_e, launches_e, clocks_e, stats_acc, all_poses = \
        timed_loop(from_host=True, fetch_poses=True)
    # ---- what the host side can feed at this N (all ranks copying at once) ----
    ceiling_gbs = h2d_ceiling(torch, dist, dev, host, world)

    # algorithmic bytes of the estimate kernel: 10 B x points x residual sweeps (both loops
    # process the same frames, so the sweeps counted in the e2e loop hold for the value loop)
    point_evals = 0
    sweeps = 0
    for out, st in stats_acc:
        for s in st:
            for lvl in range(5):
                point_evals += int(s.n_points[lvl]) * int(s.evaluations[lvl])
                sweeps += int(s.evaluations[lvl])
    est_ms, est_launches = prof["estimate"]
    peak, peak_src = measured_peak()
    # one estimate CALL per step = one launch of the dominant kernel (the dataflow form adds a
    # ~2 us ring-initialisation kernel to the same timed span; it is part of the figure)
    est_calls = max(K, 1)
    bytes_per_launch = BYTES_PER_POINT * point_evals / est_calls
    dur_s = est_ms * 1e-3 / est_calls
    achieved = bytes_per_launch / dur_s / 1e9 if dur_s > 0 else 0.0
    kernels = {}
    total_k_ms = sum(v[0] for v in prof.values())
    for k, (kms, kl) in prof.items():
        kernels[k] = {"ms_per_step": kms / K, "launches_per_step": kl / K,
                      "share": kms / total_k_ms if total_k_ms else None}
    # image kernels: bytes by SURVEY.md 8-d's formulas and bytes this design actually moves
    n0 = w * h
    sum_n = sum((w >> l) * (h >> l) for l in range(5))
    lv = (sum_n - n0) if args.lazy_levels else sum_n   # levels K2 / K3 touch
    survey = {"pyramid": B * (n0 + (sum_n - n0)), "gradient": B * (sum_n + 5 * sum_n)}
    moved = {"pyramid": B * (n0 + sum_n),      # read the new frame, write all 5 levels
             "gradient": B * 2 * lv}           # read the image, write the u8 gradient image
    for k in survey:
        if prof[k][0] > 0:
            t_s = prof[k][0] * 1e-3 / K
            kernels[k]["survey_formula_gbs"] = survey[k] / t_s / 1e9
            kernels[k]["moved_gbs"] = moved[k] / t_s / 1e9
            kernels[k]["frac_of_hbm_peak"] = kernels[k]["moved_gbs"] / peak
    kernels["estimate"]["us_per_gn_sweep_per_problem"] = \
        1e3 * est_ms * B / max(sweeps, 1) if sweeps else None

    # ---- per-rank attribution (the job's time is the MAX over ranks) ----
    mine = {"rank": rank, "ms_per_step": ms_own / K, "e2e_ms_per_step": ms_e_own / K,
            "kernel_ms_per_step": {k: v[0] / K for k, v in prof.items()},
            "sweeps_per_track": sweeps / max(len(stats_acc) * B, 1),
            "point_evals_per_step": point_evals / max(K, 1),
            "sm_mhz": clocks["sm_mhz"], "e2e_sm_mhz": clocks_e["sm_mhz"],
            "clock_reasons": sorted(set(clocks["reasons"]) | set(clocks_e["reasons"])),
            "h2d_ceiling_gbs": ceiling_gbs,
            "e2e_h2d_gbs": B * n0 / (ms_e_own / K * 1e-3) / 1e9}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---- single-problem latency: one pair alone on the GPU (16-CTA cluster) ----
    prime()
    upload(1, slots_b, False)
    t.EstimatePose([0], [B])
    t.profile(True)
    reps1 = 20
    for _ in range(reps1):
        _, st1 = t.EstimatePose([0], [B], return_stats=True)
    prof1 = t.profile_read()
    t.profile(False)
    sweeps1 = sum(st1[0].evaluations)
    gn_us = {"single_problem_latency_us_per_sweep": 1e3 * prof1["estimate"][0] / (reps1 * sweeps1),
             "single_problem_estimate_us": 1e3 * prof1["estimate"][0] / reps1,
             "batched_us_per_sweep_per_problem": 1e3 * est_ms * B / max(sweeps, 1),
             "batched_amortised_us_per_sweep": 1e3 * est_ms / max(sweeps, 1),
             "sweeps_per_track": sweeps / max(len(stats_acc) * B, 1)}

    line = None
    if rank == 0:
        value = world * B * K / (ms * 1e-3)
        e2e_value = world * B * K / (ms_e * 1e-3)
        ceil_min = min(p["h2d_ceiling_gbs"] for p in per_rank)
        e2e_gbs = B * n0 / (ms_e / K * 1e-3) / 1e9
        cluster_form = args.cluster_kernel or args.cluster or args.dmma_accum
        line = {
            "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64acc", "data": "synthetic",
            "config": workload_config(world, B, K, W),
            "impl_config": {"cluster_size": args.cluster,
                            "estimate_kernel": "cluster" if cluster_form else "dataflow",
                            "weights": ["identity", "tukey_mad", "huber"][args.weights],
                            "lazy_levels": bool(args.lazy_levels),
                            "host_numa_node_of_rank0": numa_node},
            "e2e": {"value": e2e_value, "unit": "tracks/s", "ms_per_step": ms_e / K,
                    "h2d_bytes_per_step": B * n0, "d2h_bytes_per_step":
                    B * (7 * 4 + 4 * 4 * 7),
                    "h2d_gbs_per_gpu": e2e_gbs,
                    "h2d_ceiling_gbs_per_gpu": ceil_min,
                    "h2d_ceiling_gbs_aggregate": sum(p["h2d_ceiling_gbs"] for p in per_rank),
                    "frac_of_h2d_ceiling": e2e_gbs / ceil_min if ceil_min else None,
                    "note": "ceiling = all %d ranks copying one step's pinned frames with "
                            "plain cudaMemcpyAsync at the same time (slowest rank)" % world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": ("estimate_kernel (cluster)" if cluster_form
                                    else "estimate_flow_kernel"),
                         "bound": "hbm", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(B),
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": est_ms / est_calls,
                         "point_evals_per_launch": point_evals / est_calls,
                         "real_bound": issue_bound(point_evals / est_calls, est_ms / est_calls,
                                                   clocks["sm_mhz"])},
            "kernels": kernels,
            "per_rank": per_rank,
            "gn_iteration_us": gn_us,
            "wall_s": {"value_loop": wall, "e2e_loop": wall_e},
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded CPU sample of the same workload: the first `nseq` sequences, all W+K steps,
            # one thread (the reference is single-threaded)
            nseq = min(B, args.cpu_sequences)
            fr = host[:1 + W + K, :nseq].numpy()
            trk = CpuTrackers(fr[0], 1, weight_mode=args.weights)
            dt, bad, cmp_n = 0.0, 0, 0
            for i in range(1, 1 + W + K):
                poses, d = trk.step(fr[i])
                dt += d
                gpu = all_poses[i - 1]
                for s in range(nseq):
                    cmp_n += 1
                    bad += 0 if np.array_equal(poses[s], gpu[s]) else 1
            trk.close()
            n = nseq * (W + K)
            line["cpu_baseline"] = {
                "value": n / dt, "unit": "tracks/s", "cores": 1, "kind": "port",
                "sample": "%d sequences x %d tracked frames of this workload (%d tracks, %.1f s), "
                          "single-threaded oracle" % (nseq, W + K, n, dt),
                "gpu_pose_bit_identical_on_sample": bad == 0}
            line["parity_sample"] = {"tracks_compared": cmp_n, "pose_mismatches": bad,
                                     "what": "every pose of every warm-up and timed e2e step of "
                                             "the sampled sequences, GPU vs oracle, bit for bit",
                                     "soak": soak_record()}
    t.close()
    del frames, host
    torch.cuda.empty_cache()
    # ---- BASELINE configs 1 / 3 / 4 in the same record ----
    if not args.no_secondary:
        sec = run_secondary(args, torch, dist, rank, local_rank, world)
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def issue_bound(points_per_launch, launch_ms, sm_mhz):
    """K4 against the roof that actually bounds it.  Per point the sweep issues
    K4_INSTR_PER_POINT SASS instructions (profiles/: per-pipe histogram of the hot loop), of
    which K4_FP64_PER_POINT go to the fp64 pipe; an SM issues 4 warp instructions per clock
    (128 thread-instructions) and retires 42.3 DFMA lanes per clock
    (profiles/r01_microbench_fp64.txt)."""
    try:
        with open(os.path.join(ROOT, "profiles", "k4_instruction_mix.json")) as f:
            mix = json.load(f)
    except Exception:
        return None
    clk = (sm_mhz or 1965.0) * 1e6
    sms = 148
    t_issue = points_per_launch * mix["instructions_per_point"] / (128.0 * sms * clk)
    t_fp64 = points_per_launch * (mix["dfma_per_point"] / 42.3 +
                                  mix["dadd_dmul_per_point"] / 63.6) / (sms * clk)
    return {"bound": "issue", "issue_bound_ms": 1e3 * t_issue, "fp64_pipe_bound_ms": 1e3 * t_fp64,
            "frac_of_issue_bound": 1e3 * t_issue / launch_ms if launch_ms else None,
            "frac_of_fp64_bound": 1e3 * t_fp64 / launch_ms if launch_ms else None,
            "instructions_per_point": mix["instructions_per_point"],
            "source": "profiles/k4_instruction_mix.json"}


def soak_record():
    """The committed result of the long parity soak (tools/parity_soak.py, run on a B200 box);
    bench.py only cites it."""
    try:
        with open(os.path.join(ROOT, "profiles", "parity_soak.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_secondary(args, torch, dist, rank, local_rank, world):
    """BASELINE configs 1 (EuRoC single stream), 3 (8192 pairs, strong split over the ranks) and
    4 (one 3840x2160 pair sharded over the ranks) in the driver-visible record.  Every entry is
    computed by tools/bench_configs.py's functions; a failure becomes {"error": ...} instead of
    losing the headline line."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    out = {}
    for name, fn, ns in (
            ("config1_euroc_single_stream", bc.euroc_seq, bc.Args()),
            ("config3_batch8192_640x480", bc.batch8192, bc.Args(pairs=args.secondary_pairs)),
            ("config4_shard4k", bc.shard4k, bc.Args(reps=20, check=True))):
        try:
            res = fn(ns, torch, dist, rank, local_rank, world)
        except Exception as e:  # noqa: BLE001
            res = {"error": "%s: %s" % (type(e).__name__, e)}
        out[name] = res
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--cluster", type=int, default=0, help="CTAs per problem (0 = auto)")
    ap.add_argument("--dmma-accum", action="store_true",
                    help="A/B: Gram accumulator in fp64 DMMA fragments (slower; off by default)")
    ap.add_argument("--cluster-kernel", action="store_true",
                    help="A/B: one cluster per problem instead of the dataflow kernel")
    ap.add_argument("--lazy-levels", action="store_true",
                    help="opt-in UWT_FLAG_LAZY_LEVELS: gradient/candidates only on the optimised "
                         "levels (the default does all levels, like the reference)")
    ap.add_argument("--weights", type=int, default=0,
                    help="residual weights: 0 identity (reference), 1 Tukey/MAD, 2 Huber")
    ap.add_argument("--cpu-sequences", type=int, default=24,
                    help="sequences of the workload the single-threaded CPU baseline tracks (~12 s)")
    ap.add_argument("--cpu-sequences-1t", type=int, default=8,
                    help="--impl reference: sequences of the single-thread sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip BASELINE configs 1/3/4 (the `secondary` block)")
    ap.add_argument("--secondary-pairs", type=int, default=8192)
    args = ap.parse_args()
    args.out = claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
