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
`--impl reference`: the CPU oracle (restatement of the reference, oracle/) on all host cores.

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


def cpu_tracks(frames_np, threads, accum_mode=0, weight_mode=0):
    """Runs the oracle's reference-shaped loop on [n_seq, n_frames, H, W] host frames with
    `threads` concurrent single-threaded trackers.  Returns (tracks, seconds, poses)."""
    from oracle import uw_oracle as O
    O.build()
    w, h, fx, fy, cx, cy = synth.CALIB[CALIB]
    nseq = frames_np.shape[0]
    poses = [None] * nseq
    O.lib()

    def work(ids):
        p = O.default_params(w, h, fx, fy, cx, cy, accum_mode=accum_mode, threads=1,
                             weight_mode=weight_mode)
        for s in ids:
            poses[s], _, _ = O.track_sequence(p, frames_np[s])

    groups = [list(range(i, nseq, threads)) for i in range(threads)]
    ths = [threading.Thread(target=work, args=(g,)) for g in groups if g]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return nseq * (frames_np.shape[1] - 1), dt, poses


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    dev = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    per_step = 2 * threads  # bounded sample: two tracks per host thread per step
    n_steps = args.warmup + args.steps
    # every step tracks `per_step` fresh frame pairs (sequence length 2)
    seeds = list(range(10_000, 10_000 + per_step))
    fr = gen_sequences(torch, dev, seeds, 2).permute(1, 0, 2, 3).contiguous().cpu().numpy()
    times = []
    for s in range(n_steps):
        n, dt, _ = cpu_tracks(fr, threads)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64acc", "data": "synthetic",
        "config": {"workload": "tum_mono_1280x1024_seq", "tracks_per_step": per_step,
                   "levels": 5, "optimised_levels": "4..1"},
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": threads, "kind": "port",
                         "sample": "%d tracks/step x %d steps, one single-threaded oracle "
                                   "tracker per host thread" % (per_step, args.steps)},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


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
    seeds = [rank * 100_000 + i for i in range(B)]
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
        stats_acc = []
        upload(1, cur, from_host)
        for i in range(1, 1 + W):
            track(prev, cur)
            upload(i + 1, prev, from_host)   # next step's frames go where `prev` lives
            if fetch_poses:
                fetch()
            prev, cur = cur, prev
        barrier()
        t.profile(True)
        l0 = t.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        e0.record(stream)
        for i in range(1 + W, 1 + W + K):
            track(prev, cur)
            upload(i + 1, prev, from_host)
            if fetch_poses:
                stats_acc.append(fetch())
            prev, cur = cur, prev
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        prof = t.profile_read()
        launches = t.launch_count() - l0
        t.profile(False)
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, wall, prof, launches, clocks, stats_acc

    # ---- value: inputs resident in HBM ----
    ms, wall, prof, launches, clocks, _ = timed_loop(from_host=False, fetch_poses=False)
    # ---- e2e: pinned host frames in, poses out, every step ----
    ms_e, wall_e, prof_e, launches_e, clocks_e, stats_acc = timed_loop(from_host=True, fetch_poses=True)

    # algorithmic bytes of the estimate kernel: 10 B x points x residual sweeps (both loops
    # process the same frames, so the sweeps counted in the e2e loop hold for the value loop)
    point_evals = 0
    evals_hist = np.zeros(8, np.int64)
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
        line = {
            "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64acc", "data": "synthetic",
            "config": {"workload": "tum_mono_1280x1024_seq", "sequences_per_gpu": B,
                       "tracks_per_step": world * B, "levels": 5, "optimised_levels": "4..1",
                       "candidates": "all pixels with g > mean+20",
                       "l2_policy": "inputs larger than L2 (%.0f MB of new frames + %.1f GB "
                                    "working set per step)" % (B * n0 / 1e6,
                                                                 2 * B * 21e6 / 1e9),
                       "cluster_size": args.cluster,
                       "estimate_kernel": ("cluster" if (args.cluster_kernel or args.cluster or
                                                         args.dmma_accum)
                                           else "dataflow"),
                       "weights": ["identity", "tukey_mad", "huber"][args.weights],
                       "lazy_levels": bool(args.lazy_levels),
                       "host_numa_node_of_rank0": numa_node,
                       "parallelism": "independent sequences, %d per GPU, no comms" % B},
            "e2e": {"value": e2e_value, "unit": "tracks/s", "ms_per_step": ms_e / K,
                    "h2d_bytes_per_step": B * n0, "d2h_bytes_per_step":
                    B * (7 * 4 + 4 * 4 * 7)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": ("estimate_kernel (cluster)"
                                    if (args.cluster_kernel or args.cluster or args.dmma_accum)
                                    else "estimate_flow_kernel"),
                         "bound": "hbm", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(B),
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": est_ms / est_calls,
                         "point_evals_per_launch": point_evals / est_calls},
            "kernels": kernels,
            "gn_iteration_us": gn_us,
            "wall_s": {"value_loop": wall, "e2e_loop": wall_e},
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded CPU sample of the same workload: the first `nseq` sequences, all K+W steps
            nseq = min(B, args.cpu_sequences)
            fr = host[:1 + W + K, :nseq].permute(1, 0, 2, 3).contiguous().numpy()
            n, dt, poses = cpu_tracks(fr, 1, weight_mode=args.weights)
            line["cpu_baseline"] = {
                "value": n / dt, "unit": "tracks/s", "cores": 1, "kind": "port",
                "sample": "%d sequences x %d frames of this workload (%d tracks, %.1f s), "
                          "single-threaded oracle" % (nseq, 1 + W + K, n, dt)}
            # bonus: the oracle's last pose of every sampled sequence vs the GPU's e2e result
            gpu_last = stats_acc[-1][0]
            same = all(np.array_equal(poses[s][-1], gpu_last[s]) for s in range(nseq))
            line["cpu_baseline"]["gpu_pose_bit_identical_on_sample"] = bool(same)
        print(json.dumps(line), file=args.out, flush=True)
    t.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="sequences per GPU")
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.out = claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
