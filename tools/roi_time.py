import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import jdet_b200.ops as ops
from _inputs import dota_boxes
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
feat = torch.randn((1, 256, 256, 256), device=dev)
rois = torch.as_tensor(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048)], 1)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for thr in sys.argv[1:]:
    os.environ["JDET_ROI_THREADS"] = thr
    fn = lambda: ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2)
    for _ in range(5): flush.zero_(); fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    print("threads", thr, "median us %.1f  min %.1f" % (t[25] * 1e3, t[0] * 1e3))
