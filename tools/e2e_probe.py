import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import uw_slam_b200 as U
from uw_slam_b200 import synth, _lib as L
import bench
B = int(sys.argv[1]); K = 8
dev = torch.device('cuda', 0)
w, h, fx, fy, cx, cy = synth.CALIB[bench.CALIB]
fr = bench.gen_sequences(torch, dev, list(range(B)), K + 3)
host = torch.empty(fr.shape, dtype=torch.uint8, pin_memory=True); host.copy_(fr); torch.cuda.synchronize()
t = U.Tracker(False)
t.InitializePyramid(w, h, U.CameraModel.from_intrinsics(w, h, fx, fy, cx, cy).GetK(), max_frames=2 * B)
a, b = list(range(B)), list(range(B, 2 * B))
t.AddFramesDevice(a, fr[0].data_ptr()); t.ApplyGradient(a); t.ObtainCandidatePoints(a)
prev, cur = a, b
fb = w * h
t.AddFramesHostPtr(cur, host[1].data_ptr(), w, fb)
def fetch():
    st = (L.TrackStats * B)(); out = np.empty((B, 7), np.float32)
    t._check(t._lib.uwt_fetch_poses(t._h, B, out.ctypes.data_as(L._fp), st))
rows = []
for i in range(1, 1 + K):
    t0 = time.perf_counter(); t.EstimatePoseAsync(prev, cur)
    t1 = time.perf_counter(); t.ApplyGradient(cur); t.ObtainCandidatePoints(cur)
    t2 = time.perf_counter(); t.AddFramesHostPtr(prev, host[i + 1].data_ptr(), w, fb)
    t3 = time.perf_counter(); fetch()
    t4 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
    prev, cur = cur, prev
t.synchronize()
for r in rows: print('B=%d est %.2f  k3 %.2f  upload %.2f  fetch %.2f  total %.2f ms' % ((B,) + tuple(1e3 * x for x in r)))
