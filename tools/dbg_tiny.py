import sys; sys.path.insert(0,'.')
import numpy as np
import uw_slam_b200 as U
from uw_slam_b200 import synth, _lib as L
gold=np.load('tests/golden/golden_small.npz')
key='tiny_0'
w,h,fx,fy,cx,cy=synth.CALIB['tiny']
t=U.Tracker(False); t.InitializePyramid(w,h,U.CameraModel.from_intrinsics(w,h,fx,fy,cx,cy).GetK(),max_frames=2,flags=L.FLAG_TRACE)
fp,fc=t.AddFrames([0,1],np.stack([gold[key+'_prev'],gold[key+'_cur']])); t.ApplyGradient(fp); t.ObtainCandidatePoints(fp)
pose=t.EstimatePose(fp,fc)[0]
tr=t.get_trace(0)
A=np.array([x.A[:] for x in tr],np.float32); G=gold[key+'_m0_A']
for i in range(len(tr)):
    d=np.nonzero(A[i]!=G[i])[0]
    print(i,'lvl',tr[i].level,'k',tr[i].k,'nvalid',tr[i].n_valid,'ndiff',len(d), [(int(j),float(A[i][j]),float(G[i][j])) for j in d[:4]])
    print('   pose before?', np.array(tr[i].pose[:]))
