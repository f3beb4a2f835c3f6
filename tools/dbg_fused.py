import sys; sys.path.insert(0, '/root/repo')
import numpy as np
import uw_slam_b200 as U
from uw_slam_b200 import synth
from oracle import uw_oracle as O
for (w,h,levels) in [(320,256,1),(320,256,5),(64,48,5),(1280,1024,5)]:
    rng=np.random.default_rng(1)
    img=rng.integers(0,256,(h,w),dtype=np.uint8)
    t=U.Tracker(False)
    t.InitializePyramid(w,h,U.CameraModel.from_intrinsics(w,h,500.,500.,w/2-.5,h/2-.5).GetK(),max_frames=2,levels=levels,first_level=levels-1,last_level=min(1,levels-1))
    f=t.AddFrames([0],img)[0]
    ref=O.FrameData(img,levels=levels)
    for l in range(levels):
        g=t.get_gradient_image(0,l)
        im=t.get_image(0,l)
        bad=np.argwhere(g!=ref.g[l])
        print(w,h,levels,"lvl",l,"img ok",np.array_equal(im,ref.images[l]),"g mismatches",len(bad), "first", bad[:6].tolist(), "rows", sorted(set(bad[:,0].tolist()))[:8], "cols", sorted(set(bad[:,1].tolist()))[:8])
    t.ObtainCandidatePoints([0])
    for l in range(levels):
        c=t.get_candidates(0,l)
        print("   cand lvl",l,c.shape, ref.cand[l].shape, np.array_equal(c,ref.cand[l]), "ithr ref", ref.ithr[l])
    t.close()
