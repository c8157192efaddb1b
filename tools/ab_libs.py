#!/usr/bin/env python
"""A/B two or more builds of libjdet_b200.so on one GPU box: same inputs, CUDA-event timing, result hashes.

  python tools/ab_libs.py build  NAME=git:HEAD NAME2=flags:-DJDET_SAFE_DIV ...   # on the build host: tools/_build/ab/libNAME.so
  python tools/ab_libs.py run [--ops iou,nms,roi] [NAME ...]                    # on the GPU box: one subprocess per library

`run` prints one JSON line per library: microseconds per call (L2 flushed between calls) and a hash of every result,
so a faster build that changes a single bit of an IoU matrix or one NMS keep flag is visible immediately.
The working-tree build (jdet_b200/_C/libjdet_b200.so) is always included as "tree".
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "_build", "ab")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(specs):
    from jdet_b200 import _lib
    os.makedirs(AB, exist_ok=True)
    for spec in specs:
        name, what = spec.split("=", 1)
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "csrc")
            os.makedirs(src)
            flags = []
            if what.startswith("git:"):
                rev = what[4:]
                for f in subprocess.check_output(["git", "-C", ROOT, "ls-tree", "--name-only", rev, "jdet_b200/csrc/"]).decode().split():
                    open(os.path.join(src, os.path.basename(f)), "wb").write(subprocess.check_output(["git", "-C", ROOT, "show", "%s:%s" % (rev, f)]))
            else:
                assert what.startswith("flags:")
                flags = what[6:].split()
                for f in os.listdir(_lib.CSRC):
                    open(os.path.join(src, f), "wb").write(open(os.path.join(_lib.CSRC, f), "rb").read())
            objs, procs = [], []
            for f in sorted(os.listdir(src)):
                if f.endswith(".cu"):
                    o = os.path.join(tmp, f[:-3] + ".o")
                    objs.append(o)
                    procs.append(subprocess.Popen(["nvcc"] + _lib.NVCC_FLAGS + flags + ["-c", os.path.join(src, f), "-o", o]))
            assert all(p.wait() == 0 for p in procs)
            out = os.path.join(AB, "lib%s.so" % name)
            subprocess.check_call(["nvcc", "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
            print("built", out)


def worker(ops_sel):
    import numpy as np
    import torch
    import jdet_b200.ops as ops
    from _inputs import clustered_boxes, dota_boxes, tie_free_scores
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    cu = lambda x, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(x), dtype=dt).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h = lambda t: hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()[:12]

    def timeit(fn, k):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(k):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return round(ts[len(ts) // 2], 1)

    res = {"lib": os.environ.get("JDET_B200_LIB", "tree")}
    if "iou" in ops_sel:
        b1, b2 = cu(dota_boxes(rng, 16384)), cu(dota_boxes(rng, 16384))
        res["iou16k_us"] = timeit(lambda: ops.box_iou_rotated(b1, b2), 10)
        res["iou16k_hash"] = h(ops.box_iou_rotated(b1, b2))
        c = cu(clustered_boxes(rng, 4096, 50))
        res["iou4k_clustered_us"] = timeit(lambda: ops.box_iou_rotated(c, c), 10)
        res["iou4k_clustered_hash"] = h(ops.box_iou_rotated(c, c)) + "/" + h(ops.box_iou_rotated_v1(c, c))
        s1, s2 = cu(dota_boxes(rng, 1000)), cu(dota_boxes(rng, 1000))
        res["iou1k_us"] = timeit(lambda: ops.box_iou_rotated(s1, s2), 30)
        res["iou1k_hash"] = h(ops.box_iou_rotated(s1, s2))
    if "nms" in ops_sel:
        n = 100000
        d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
        td, ts_, tl = cu(d), cu(tie_free_scores(rng, n)), cu(rng.integers(0, 15, n), torch.int64)
        for thr in (0.1, 0.5):
            res["nms100k_thr%.1f_us" % thr] = timeit(lambda: ops.nms_rotated.ml_nms_rotated(td, ts_, tl, thr), 10)
            res["nms100k_thr%.1f_hash" % thr] = h(ops.nms_rotated.ml_nms_rotated(td, ts_, tl, thr))
        res["nms100k_agnostic_us"] = timeit(lambda: ops.nms_rotated.nms_rotated(td[:30000], ts_[:30000], 0.3), 5)
        res["nms100k_agnostic_hash"] = h(ops.nms_rotated.nms_rotated(td[:30000], ts_[:30000], 0.3))
    if "roi" in ops_sel:
        feat = torch.randn((1, 256, 256, 256), device=dev, generator=torch.Generator(dev).manual_seed(1))
        rois = cu(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048, 1024.0)], 1))
        res["roi_cfg2_us"] = timeit(lambda: ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2), 30)
        res["roi_cfg2_hash"] = h(ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2))
        # the same RoIs handed over largest first / smallest first: how much of the time is the tail of the work queue?
        area = rois[:, 3] * rois[:, 4]
        big = rois[torch.argsort(area, descending=True)].contiguous()
        small = rois[torch.argsort(area)].contiguous()
        res["roi_cfg2_largest_first_us"] = timeit(lambda: ops.roi_align_rotated_v1.roi_align(feat, big, (7, 7), 0.25, 2), 30)
        res["roi_cfg2_smallest_first_us"] = timeit(lambda: ops.roi_align_rotated_v1.roi_align(feat, small, (7, 7), 0.25, 2), 30)
        fcl = feat.contiguous(memory_format=torch.channels_last)
        res["roi_cfg2_channels_last_us"] = timeit(lambda: ops.roi_align_rotated_v1.roi_align(fcl, rois, (7, 7), 0.25, 2), 30)
        res["roi_cfg2_channels_last_largest_first_us"] = timeit(lambda: ops.roi_align_rotated_v1.roi_align(fcl, big, (7, 7), 0.25, 2), 30)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "worker":
        worker(sys.argv[2].split(","))
    else:
        args = sys.argv[2:]
        ops_sel = "iou,nms,roi"
        if args and args[0] == "--ops":
            ops_sel, args = args[1], args[2:]
        names = args or (sorted(f[3:-3] for f in os.listdir(AB) if f.endswith(".so")) if os.path.isdir(AB) else [])
        order = names if "tree" in names else ["tree"] + names      # "tree" may be placed explicitly (first process on a cold box runs slow)
        for name in order:
            env = dict(os.environ)
            env.pop("JDET_B200_LIB", None)
            if name != "tree":
                env["JDET_B200_LIB"] = os.path.join(AB, "lib%s.so" % name)
            subprocess.call([sys.executable, os.path.abspath(__file__), "worker", ops_sel], env=env)
