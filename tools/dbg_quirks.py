import sys; sys.path.insert(0,'/root/repo')
import numpy as np, scheme_raytrace_b200 as srt
w=h=48
r = srt.Renderer(srt.scenes.cfg4_cornell_box(w,h), lights=[2])
for q in (15, 14, 0, 1):
    for est in (0,1):
        a,_ = r.render(w,h,256,seed=12,quirks=q,estimator=est)
        bad = np.argwhere(~np.isfinite(a) | (a<0))
        print('quirks',q,'est',est,'mean',a.mean()/256,'min',a.min(),'nbad',len(bad), bad[:3].tolist())
