#!/usr/bin/env python
"""Parity soak: many 1280x1024 tracks on the batched dataflow kernel, EVERY Gauss-Newton sweep's
N_valid, sum r^2, error, A (6x6), b, delta and pose compared bit for bit with the CPU oracle.

The normal equations are fp64 sums rounded once to f32 (docs/ARITHMETIC.md U3); the oracle adds
sequentially, the GPU thread -> warp tree -> chunk order.  The two can differ only when an fp64
sum lands within its own rounding error of an f32 rounding boundary; this tool measures how
often that happens.  Pairs are seeded; a share of them gets occluders (a rectangle of the
current frame replaced) or is made sparse (large flat regions, few candidates).

    python tools/parity_soak.py --tracks 100000 --out gpurun_out/parity_soak.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from uw_slam_b200 import synth  # noqa: E402

TRACE_FIELDS = ("level", "k", "n_valid", "broke", "sum_r2", "error", "A", "b", "delta", "pose")


def trace_rows(tr):
    """ctypes trace records -> list of byte strings per field (bit-exact comparison)."""
    rows = []
    for r in tr:
        rows.append((r.level, r.k, r.n_valid, r.broke, r.sum_r2,
                     np.float32(r.error).tobytes(), bytes(r.A), bytes(r.b), bytes(r.delta),
                     bytes(r.pose)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=100000)
    ap.add_argument("--calib", default="tum_mono")
    ap.add_argument("--batch", type=int, default=64, help="problems per call (<= 64: traced)")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--seed0", type=int, default=1_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_soak.json"))
    args = ap.parse_args()
    import torch
    import uw_slam_b200 as U
    from uw_slam_b200 import _lib as L
    from oracle import uw_oracle as O
    O.build()
    O.lib()
    dev = torch.device("cuda", 0)
    w, h, fx, fy, cx, cy = synth.CALIB[args.calib]
    B = args.batch
    t = U.Tracker(False)
    # more than 64 problems per call are not traced by the library: such a soak exercises the
    # large-batch task partition (16384-record chunks from 192 problems on) and compares the final
    # poses only
    traced = B <= 64
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(),
                        max_frames=2 * B, flags=L.FLAG_TRACE if traced else 0)
    ps, cs = list(range(B)), list(range(B, 2 * B))
    params = [O.default_params(w, h, fx, fy, cx, cy, threads=1) for _ in range(args.threads)]
    pool = ThreadPoolExecutor(max_workers=args.threads)
    rng = np.random.default_rng(12345)
    ys = torch.arange(h, device=dev)[:, None]
    xs = torch.arange(w, device=dev)[None, :]

    tot = {"tracks": 0, "sweeps": 0, "sweep_mismatches": 0, "pose_mismatches": 0,
           "field_mismatches": {f: 0 for f in TRACE_FIELDS}, "trace_length_mismatches": 0,
           "variants": {"plain": 0, "occluder": 0, "sparse": 0}, "examples": []}
    t_start = time.time()
    done = 0
    while done < args.tracks:
        seeds = [args.seed0 + done + i for i in range(B)]
        # motions up to 1.5x the headline magnitude so that sweep counts vary
        scale = float(rng.uniform(0.3, 1.5))
        prev, cur = synth.render_pairs_torch(args.calib, seeds, dev, rot=5e-3 * scale,
                                             trans=5e-3 * scale)
        kinds = rng.choice(3, size=B, p=[0.6, 0.2, 0.2])
        for i, kd in enumerate(kinds):
            if kd == 1:    # occluder: a rectangle of the current frame replaced by noise / a flat patch
                x0, y0 = int(rng.integers(0, w - 200)), int(rng.integers(0, h - 200))
                ww, hh = int(rng.integers(40, 400)), int(rng.integers(40, 400))
                if rng.random() < 0.5:
                    cur[i, y0:y0 + hh, x0:x0 + ww] = int(rng.integers(0, 256))
                else:
                    cur[i, y0:y0 + hh, x0:x0 + ww] = torch.randint(
                        0, 256, cur[i, y0:y0 + hh, x0:x0 + ww].shape, device=dev, dtype=torch.uint8)
                tot["variants"]["occluder"] += 1
            elif kd == 2:  # sparse: keep texture only inside a few blobs, flat grey elsewhere
                keep = torch.zeros((h, w), dtype=torch.bool, device=dev)
                for _ in range(int(rng.integers(1, 5))):
                    bx, by = int(rng.integers(0, w)), int(rng.integers(0, h))
                    r = int(rng.integers(30, 250))
                    keep |= ((xs - bx) ** 2 + (ys - by) ** 2) < r * r
                prev[i] = torch.where(keep, prev[i], torch.full_like(prev[i], 128))
                cur[i] = torch.where(keep, cur[i], torch.full_like(cur[i], 128))
                tot["variants"]["sparse"] += 1
            else:
                tot["variants"]["plain"] += 1
        torch.cuda.synchronize()   # torch rendered on its own stream; the library has its own
        t.AddFramesDevice(ps, prev.data_ptr())
        t.AddFramesDevice(cs, cur.data_ptr())
        t.ApplyGradient(ps)
        t.ObtainCandidatePoints(ps)
        gposes = t.EstimatePose(ps, cs)
        gtr = [trace_rows(t.get_trace(i)) if traced else None for i in range(B)]
        hp, hc = prev.cpu().numpy(), cur.cpu().numpy()

        def cpu(job):
            tid, idx = job
            out = []
            for i in idx:
                op, _, otr = O.estimate_pose(params[tid], O.FrameData(hp[i]),
                                             O.FrameData(hc[i], with_candidates=False))
                out.append((i, op, trace_rows(otr)))
            return out
        jobs = [(k, list(range(k, B, args.threads))) for k in range(args.threads)]
        for res in pool.map(cpu, jobs):
            for i, op, otr in res:
                tot["tracks"] += 1
                if not np.array_equal(op, gposes[i]):
                    tot["pose_mismatches"] += 1
                if not traced:
                    tot["sweeps"] += len(otr)
                    continue
                if len(otr) != len(gtr[i]):
                    tot["trace_length_mismatches"] += 1
                for a, b in zip(otr, gtr[i]):
                    tot["sweeps"] += 1
                    bad = [f for f, x, y in zip(TRACE_FIELDS, a, b) if x != y]
                    if bad:
                        tot["sweep_mismatches"] += 1
                        for f in bad:
                            tot["field_mismatches"][f] += 1
                        if len(tot["examples"]) < 8:
                            ex = {"seed": seeds[i], "variant": int(kinds[i]), "level": a[0],
                                  "k": a[1], "fields": bad}
                            for f in ("A", "b"):
                                if f in bad:
                                    j = TRACE_FIELDS.index(f)
                                    xa = np.frombuffer(a[j], np.float32)
                                    xb = np.frombuffer(b[j], np.float32)
                                    d = np.nonzero(xa.view(np.int32) != xb.view(np.int32))[0]
                                    ex[f] = [{"index": int(q), "oracle": float(xa[q]),
                                              "gpu": float(xb[q]),
                                              "ulps": int(abs(int(xa.view(np.int32)[q]) -
                                                              int(xb.view(np.int32)[q])))}
                                             for q in d[:4]]
                            tot["examples"].append(ex)
        done += B
        if (done // B) % 50 == 0:
            print("[soak] %d tracks, %d sweeps, %d sweep mismatches, %d pose mismatches, %.0f s"
                  % (tot["tracks"], tot["sweeps"], tot["sweep_mismatches"],
                     tot["pose_mismatches"], time.time() - t_start), file=sys.stderr, flush=True)
    tot["seconds"] = time.time() - t_start
    tot["config"] = {"calib": args.calib, "size": [w, h], "kernel": "estimate_flow_kernel "
                     "(batches of %d%s)" % (B, ", UWT_FLAG_TRACE" if traced else ", untraced"),
                     "threads": args.threads, "seed0": args.seed0,
                     "compared": ("every sweep: level, k, N_valid, sum r^2, error, A[36], b[6], "
                                  "delta[6], pose[7]; and the final pose; bit for bit") if traced
                     else "the final pose of every track, bit for bit (sweeps = the oracle's count)"}
    tot["sweep_mismatch_rate"] = tot["sweep_mismatches"] / max(tot["sweeps"], 1)
    tot["pose_mismatch_rate"] = tot["pose_mismatches"] / max(tot["tracks"], 1)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(tot, f, indent=1)
    print(json.dumps({k: v for k, v in tot.items() if k != "examples"}))
    t.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
