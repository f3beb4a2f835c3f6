import sys, time
sys.path.insert(0, '.')
import numpy as np
import uw_slam_b200 as U
from uw_slam_b200 import synth
from uw_slam_b200.sharded import connect_fused
for calib in ['tum', 'tum_mono', 'uhd']:
    w, h, fx, fy, cx, cy = synth.CALIB[calib]
    prev, cur, _, _ = synth.render_pair(calib, 3)
    t = U.Tracker(False)
    t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(), max_frames=2)
    t.AddFrames([0, 1], np.stack([prev, cur])); t.ApplyGradient([0]); t.ObtainCandidatePoints([0])
    connect_fused(t)
    ref, st = t.EstimatePose([0], [1], return_stats=True)
    sweeps = sum(st[0].evaluations)
    for name, fn in [('cluster16', lambda: t.EstimatePose([0], [1])),
                     ('fused148', lambda: (t.ShardEstimateFusedAsync(0, 1, grid=148), t.ShardEstimateFusedWait())[1][0]),
                     ('fused296', lambda: (t.ShardEstimateFusedAsync(0, 1, grid=296), t.ShardEstimateFusedWait())[1][0]),
                     ('fused74', lambda: (t.ShardEstimateFusedAsync(0, 1, grid=74), t.ShardEstimateFusedWait())[1][0])]:
        for _ in range(5): out = fn()
        t.profile(True)
        t0 = time.perf_counter()
        for _ in range(50): out = fn()
        dt = (time.perf_counter() - t0) / 50
        ms = t.profile_read()['estimate'][0] / 50
        t.profile(False)
        same = np.array_equal(np.asarray(out).reshape(-1)[:7], ref[0])
        print(calib, name, 'wall %.1f us  kernel %.1f us  per sweep %.2f us  sweeps %d same %s' % (dt * 1e6, ms * 1e3, ms * 1e3 / sweeps, sweeps, same))
    t.close()
