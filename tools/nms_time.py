import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, time
import jdet_b200.ops as ops
from _inputs import *
rng=np.random.default_rng(100); n=100000
d=np.concatenate([clustered_boxes(rng,n//2,50),dota_boxes(rng,n-n//2)]); s=tie_free_scores(rng,n); l=rng.integers(0,15,n)
td,ts,tl=[torch.as_tensor(a).cuda() for a in (d,s,l)]
for thr in (0.1,0.5):
    for _ in range(3): ops.nms_rotated.ml_nms_rotated(td,ts,tl,thr)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): k=ops.nms_rotated.ml_nms_rotated(td,ts,tl,thr)
    torch.cuda.synchronize(); print("ml_nms 100k thr",thr,"ms", (time.time()-t)*100, k.numel())
