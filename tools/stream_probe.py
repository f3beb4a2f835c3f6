"""Where a single-stream track spends its time (config 1): host time of every API call."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import uw_slam_b200 as U
from uw_slam_b200 import synth
calib = "euroc"
w, h, fx, fy, cx, cy = synth.CALIB[calib]
frames, _, _ = synth.render_sequence(calib, 11, 64, rot=1.5e-3, trans=1.5e-3)
host = torch.from_numpy(np.stack(frames)).pin_memory()
t = U.Tracker(False)
t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(), max_frames=2)
KEY, CUR = 0, 1
for rep in range(2):
    rows = []
    t.AddFramesHostPtr([KEY], host[0].data_ptr(), w, w * h); t.ApplyGradient([KEY]); t.ObtainCandidatePoints([KEY])
    t.synchronize()
    for i in range(1, 64):
        a = time.perf_counter()
        t.AddFramesHostPtr([CUR], host[i].data_ptr(), w, w * h)
        b = time.perf_counter()
        pose = t.EstimatePose([KEY], [CUR])
        c = time.perf_counter()
        if i % 8 == 0:
            t.AddFramesHostPtr([KEY], host[i].data_ptr(), w, w * h); t.ApplyGradient([KEY]); t.ObtainCandidatePoints([KEY])
        d = time.perf_counter()
        rows.append((b - a, c - b, d - c))
    r = np.array(rows) * 1e6
print("upload call %.1f us  estimate call (sync) %.1f us  rekey (amortised) %.1f us  total %.1f us" % (r[:, 0].mean(), r[:, 1].mean(), r[:, 2].mean(), r.sum(1).mean()))
t.profile(True)
for i in range(1, 64):
    t.AddFramesHostPtr([CUR], host[i].data_ptr(), w, w * h); t.EstimatePose([KEY], [CUR])
p = t.profile_read(); print({k: (round(1e3 * v[0] / 63, 1), v[1]) for k, v in p.items()})
