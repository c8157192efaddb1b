#!/usr/bin/env python
"""bench.py — JDet oriented-box geometry hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extra]

Headline (BASELINE.json `metric`, quoted on configs[1]): rotated RoIs/s of roi_align_rotated on a
256-ch 256x256 FPN map with 2048 RoIs, 7x7 bins, sampling_ratio 2.  A "step" is one pass of that
op over one synthetic 1024^2-tile's RoI batch.  `value` is measured with inputs resident in HBM,
`e2e` through the public op with pinned HOST buffers (H2D + op + D2H inside the timed region).
The second half of the metric (rNMS boxes/s, configs[2]) and the other hot-path ops are reported in
`extra`, each with its own roofline line.  N > 1: one process per GPU (torchrun), each rank runs the
same per-GPU workload on its own tile (weak scaling, no data-path collective in the headline op);
the NMS leg all-gathers padded final detections over NCCL.

`--impl reference` times the reference's CPU implementation of the same path on the host cores:
roi_align_rotated has NO CPU path in the reference (CUDA-only), so this arm runs the oracle port
(`kind: port`, OpenMP over RoIs) on a bounded RoI sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

L2_FLUSH_BYTES = 256 << 20
PREWARM_S = 1.5
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} \
            if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
            {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def time_steps(torch, fn, steps, warmup, flush, sampler=None):
    """W warm-ups, then K steps; every step is bracketed by CUDA events on the current stream, the
    L2 flush sits OUTSIDE the events.  Returns total device milliseconds over the K steps."""
    for _ in range(warmup):
        flush()
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        for a, b in ev:
            flush()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
    return float(sum(a.elapsed_time(b) for a, b in ev))


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


# ------------------------------------------------------------------------------------------------
def make_cfg2(seed):
    from _inputs import dota_boxes
    rng = np.random.default_rng(seed)
    rois = np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048, 1024.0)], 1)
    return rois


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)      # before any pinned allocation: host buffers land on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import jdet_b200.ops as ops
    from jdet_b200 import dist as jdist
    peaks = load_peaks()
    flush_buf = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    clean_buf = torch.zeros(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=dev)

    def flush():
        # write 256 MiB (> 126 MB L2): nothing of the previous step survives in L2 ...
        flush_buf.zero_()
        # ... then read another 256 MiB so the dirty lines of that write are written back BEFORE the timed
        # region instead of stealing HBM bandwidth inside it
        clean_buf.max()
    K, W = args.steps, max(args.warmup, 3)

    # ---- headline: roi_align_rotated_v1, cfg2 -------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feat = torch.randn((1, 256, 256, 256), device=dev, generator=g)
    rois_h = make_cfg2(rank)
    rois = torch.as_tensor(rois_h).to(dev)
    roi_fn = lambda: ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2)
    out = roi_fn()
    torch.cuda.synchronize()
    # A box that has just been handed over runs its first seconds of GPU work well below steady state (measured with
    # tools/ab_libs.py: the first process on a fresh box 1.3-1.7x slower than the same build a few seconds later), so
    # the W warm-up steps are preceded by PREWARM_S seconds of the same untimed steps.
    t0 = time.time()
    while time.time() - t0 < PREWARM_S:
        flush()
        roi_fn()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    total_ms = time_steps(torch, roi_fn, K, W, flush, sampler)
    torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    ms_per_step = total_ms / K
    n_rois = rois.shape[0]
    value = n_rois * world / (ms_per_step * 1e-3)
    alg_bytes = feat.numel() * 4 + rois.numel() * 4 + out.numel() * 4        # 169 918 464 B
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm",
                "kernel": "roi_align_rotated = roi_prologue_kernel<1> (re-layout + tap tables + cost buckets) + roi_gather_kernel<16, true> "
                          "(2 kernel launches + one 256-B memset node per step; the duration used is the WHOLE step, dominant "
                          "kernel = the gather, ~69 % of it: 72.5 of 104.6 us in the ncu launch list, profiles/r02_launches_bench.md)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "peak_source": peaks["source"], "algorithmic_bytes": alg_bytes,
                # dram__bytes_read.sum + dram__bytes_write.sum of the two kernels from the committed ncu --set full
                # captures, cold L2.  Round 1 (profiles/r01_ncu_reentry_selected_metrics.csv): prologue 67.2 + 31.7 MB, gather
                # 114.1 + 53.0 MB = 266 MB; round 2 (profiles/r02_ncu_selected_metrics.csv): 67.2 + 0.03 and 113.6 + 0.06 MB —
                # the same reads, the written lines still in L2 when each kernel ends.  The larger figure is reported.
                "traffic": 266_000_000,
                "note": "latency-bound, no unit saturated (ncu round 2: gather l1tex 57 %, issue slots 39 %, L1 hit 57 %, 0.41 GB "
                        "L2 -> L1; prologue issue 58 %): 0.80 GB of merged taps cross L1 per launch for 170 MB of algorithmic "
                        "bytes; an L2-resident random 1-KB gather probe reaches 19-20 TB/s on this GPU "
                        "(profiles/r01_l2_gather_probe.txt)"}

    # ---- e2e: host buffers, H2D + op + D2H inside the timed region ----------------------------
    # Three streams (H2D / op / D2H), two buffers: step i+1's upload overlaps step i's op and download
    # (PCIe is full duplex).  Every step still uploads its own inputs from pinned host memory and
    # downloads its own result; the clock runs from before the first upload to after the last download.
    feat_h = feat.cpu().pin_memory()
    rois_p = torch.as_tensor(rois_h).pin_memory()
    out_h = [torch.empty(out.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    feat_d = [torch.empty_like(feat) for _ in range(2)]
    rois_d = [torch.empty_like(rois) for _ in range(2)]
    sH, sC, sD = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def e2e_run(steps):
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_op = [torch.cuda.Event() for _ in range(2)]
        ev_dn = [torch.cuda.Event() for _ in range(2)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(sH):
            t0.record()
        for i in range(steps):
            b = i & 1
            with torch.cuda.stream(sH):
                if i >= 2:
                    sH.wait_event(ev_op[b])            # the op that last read this device buffer is done
                feat_d[b].copy_(feat_h, non_blocking=True)
                rois_d[b].copy_(rois_p, non_blocking=True)
                ev_up[b].record()
            with torch.cuda.stream(sC):
                sC.wait_event(ev_up[b])
                o = ops.roi_align_rotated_v1.roi_align(feat_d[b], rois_d[b], (7, 7), 0.25, 2)
                ev_op[b].record()
            with torch.cuda.stream(sD):
                sD.wait_event(ev_op[b])
                if i >= 2:
                    sD.wait_event(ev_dn[b])
                o.record_stream(sD)
                out_h[b].copy_(o, non_blocking=True)
                ev_dn[b].record()
        with torch.cuda.stream(sD):
            t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    Ke = max(4, min(K, 40))
    e2e_run(4)
    e2e_ms = e2e_run(Ke)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": n_rois * world / (e2e_ms / Ke * 1e-3), "unit": "RoIs/s", "ms_per_step": e2e_ms / Ke, "steps": Ke,
           "h2d_bytes_per_step": int(feat.numel() * 4 + rois.numel() * 4), "d2h_bytes_per_step": int(out.numel() * 4),
           "pipeline": "3 streams, 2 buffers: upload(i+1) || op(i) || download(i-1); PCIe-bound", "host_affinity": numa}

    line = {"metric": "rotated RoIs/s", "value": value, "unit": "RoIs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rois_per_gpu": n_rois, "l2": "256 MiB memset + 256 MiB read between timed steps (outside the events): L2 holds no input and no dirty line",
                       "timing": "CUDA events per step on the launch stream, max over ranks",
                       "prewarm_s": PREWARM_S},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": 2 * K, "roofline": roofline}

    extra = {}
    if not args.no_extra:
        extra = run_extra(torch, dist if world > 1 else None, ops, jdist, dev, rank, world, peaks, flush)
        if rank == 0 and world == 1:
            extra["reference_cuda"] = reference_cuda_legs(torch, dev, flush, ms_per_step, extra)
    # compact table of every other leg, inside `config` (the driver's record keeps `config` whole but only the key names of
    # `extra`): leg -> [value, unit, ms_per_step, roofline frac]; reference_cuda: leg -> [reference kernel ms, ours ms, ratio]
    line["config"]["legs"] = {k: [float("%.6g" % v["value"]), v["unit"], round(v["ms_per_step"], 4), round(v["roofline"]["frac"], 4)]
                              for k, v in extra.items() if isinstance(v, dict) and "roofline" in v}
    if "reference_cuda" in extra:
        line["config"]["legs_vs_reference_cuda_kernels"] = {k: [round(v["reference_ms"], 4), round(v["ours_ms"], 4), round(v["ratio"], 2)]
                                                            for k, v in extra["reference_cuda"].items() if isinstance(v, dict)}
    if rank == 0 and world == 1 and not args.no_extra:
        line["cpu_baseline"] = cpu_baseline_roi(feat.cpu().numpy(), rois_h)
    line["extra"] = extra
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extra(torch, dist, ops, jdist, dev, rank, world, peaks, flush):
    """The other hot-path ops at their BASELINE configs, each with its roofline line."""
    from _inputs import clustered_boxes, dota_boxes, s2anet_anchors, tie_free_scores
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    rng = np.random.default_rng(100 + rank)
    cu = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    ex = {}

    def agg(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # cfg3: nms_rotated, 100k proposals x 15 classes, thr 0.1 (+ all-gather of padded detections at N > 1)
    n = 100000
    d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
    s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
    td, ts, tl = cu(d), cu(s), cu(l, torch.int64)

    # NMS epilogue writes the padded record straight into the persistent send buffer; one collective at N > 1
    def nms_fn():
        return jdist.nms_and_gather([(td, ts, tl)], 0.1, 2000)

    kept = ops.nms_rotated.ml_nms_rotated(td, ts, tl, 0.1)
    got = nms_fn()
    assert int(got[0, 0, -1, 0].item()) == min(int(kept.numel()), 2000)
    K = 10
    ms = agg(time_steps(torch, nms_fn, K, 3, flush)) / K
    keep_fn = lambda: ops.nms_rotated.ml_nms_rotated(td, ts, tl, 0.1)
    ms_keep = agg(time_steps(torch, keep_fn, K, 3, flush)) / K
    alg = n * (24 + 4 + 4 + 1)
    ex["nms_rotated"] = {"metric": "rNMS boxes/s", "value": n * world / (ms * 1e-3), "unit": "boxes/s", "ms_per_step": ms,
                         "steps": K, "config": {"workload": "ml_nms_rotated: 100k proposals (50% clustered) x 15 classes, IoU thr 0.1 "
                                                            "(BASELINE configs[2])", "kept": int(kept.numel()),
                                                "ms_keep_indices_only": ms_keep,
                                                "step": "argsort + NMS + pack into the send buffer" + (" + all_gather_into_tensor of 2001x7 per rank" if dist else "")},
                         "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg,
                                      "note": "latency/ALU-bound by construction (sequential greedy dependency); bytes are 3.3 MB"}}

    # box_iou_rotated 16k x 16k (1 GiB of IoUs) and cfg1 1k x 1k
    for nn, K in ((16384, 10), (1000, 50)):
        b1, b2 = cu(dota_boxes(rng, nn)), cu(dota_boxes(rng, nn))
        fn = lambda: ops.box_iou_rotated(b1, b2)
        fn()
        ms = agg(time_steps(torch, fn, K, 3, flush)) / K
        alg = 4 * nn * nn + 20 * 2 * nn
        ex["box_iou_rotated_%dk" % (nn // 1000)] = {
            "metric": "pairs/s", "value": nn * nn * world / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "steps": K,
            "config": {"workload": "box_iou_rotated %dx%d DOTA-shaped OBBs" % (nn, nn)},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg}}
        del b1, b2

    # roi_align_rotated backward at cfg2: grad_out (2048,256,7,7) -> grad_in (1,256,256,256)
    from jdet_b200.ops.roi_align_rotated_v1 import _roi_align_backward_impl
    rois_b = cu(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048, 1024.0)], 1))
    gbw = torch.Generator(device=dev).manual_seed(5 + rank)
    go = torch.randn((2048, 256, 7, 7), device=dev, generator=gbw)
    fn = lambda: _roi_align_backward_impl(1, go, rois_b, (1, 256, 256, 256), (7, 7), 0.25, 2)
    fn()
    K = 20
    ms = agg(time_steps(torch, fn, K, 3, flush)) / K
    alg = go.numel() * 4 + 256 * 256 * 256 * 4 + rois_b.numel() * 4
    ex["roi_align_rotated_backward"] = {
        "metric": "RoIs/s", "value": 2048 * world / (ms * 1e-3), "unit": "RoIs/s", "ms_per_step": ms, "steps": K,
        "config": {"workload": "ROIAlignRotated_v1 backward w.r.t. input at cfg2 (channel-last scratch + 16-B vector atomics + relayout)"},
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg}}
    del go

    # cfg2 again with the FPN map already channel-last in memory (torch.channels_last): no re-layout pass
    gcl = torch.Generator(device=dev).manual_seed(6 + rank)
    fcl = torch.randn((1, 256, 256, 256), device=dev, generator=gcl).contiguous(memory_format=torch.channels_last)
    fn = lambda: ops.roi_align_rotated_v1.roi_align(fcl, rois_b, (7, 7), 0.25, 2)
    fn()
    K = 30
    ms = agg(time_steps(torch, fn, K, 3, flush)) / K
    alg = fcl.numel() * 4 + rois_b.numel() * 4 + 2048 * 256 * 49 * 4
    ex["roi_align_rotated_channels_last"] = {
        "metric": "RoIs/s", "value": 2048 * world / (ms * 1e-3), "unit": "RoIs/s", "ms_per_step": ms, "steps": K,
        "config": {"workload": "roi_align_rotated_v1 at cfg2, input handed over as a torch.channels_last tensor (jdet_roi_align_rotated_nhwc)"},
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg}}
    del fcl

    # cfg5 (per GPU share): Oriented R-CNN RoI stage, 2 x 1024^2 tiles, 2000 proposals each through the 4-level
    # OrientedSingleRoIExtractor (rotated RoIAlign v1 on strides 4..32) + the shared FCs of OrientedHead
    from jdet_b200.models.roi_heads import OrientedHead
    gh = torch.Generator(device=dev).manual_seed(9 + rank)
    fpn = [torch.randn((2, 256, 1024 // s, 1024 // s), device=dev, generator=gh) for s in (4, 8, 16, 32)]
    ohead = OrientedHead(num_classes=15).to(dev).eval().requires_grad_(False)
    props = [torch.as_tensor(np.concatenate([dota_boxes(rng, 2000, 1024.0), rng.random((2000, 1), dtype=np.float32)], 1)).to(dev)
             for _ in range(2)]
    rois_o = ohead.arb2roi(props)
    ext = lambda: ohead.bbox_roi_extractor(fpn, rois_o)
    ext()
    K = 20
    ms_ext = agg(time_steps(torch, ext, K, 3, flush)) / K
    full = lambda: ohead(fpn, props)
    full()
    ms_full = agg(time_steps(torch, full, K, 3, flush)) / K
    alg = sum(f.numel() for f in fpn) * 4 + rois_o.numel() * 4 + 4000 * 256 * 49 * 4
    ex["oriented_rcnn_roi_stage"] = {
        "metric": "RoIs/s", "value": 4000 * world / (ms_ext * 1e-3), "unit": "RoIs/s", "ms_per_step": ms_ext, "steps": K,
        "config": {"workload": "OrientedSingleRoIExtractor: 2 images x 2000 proposals, 4 FPN levels (256 ch, strides 4-32 of 1024^2 tiles), "
                               "ROIAlignRotated_v1 7x7 sampling 2; head_ms = extractor + 2 shared FCs + cls/reg + decode + threshold",
                   "head_ms_per_step": ms_full, "images_per_s_head": 2 * world / (ms_full * 1e-3)},
        "roofline": {"bound": "hbm", "achieved": alg / (ms_ext * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": alg / (ms_ext * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg}}
    # cfg5 heads pipeline on this rank's shard (2 tiles): OrientedRPNHead (5 levels) -> 2000 proposals/img -> OrientedHead
    # -> per-class rotated NMS (the tile-level post-process) -> ONE all-gather of padded detections at N > 1.
    # Backbone + FPN are out of scope: the FPN maps are synthetic.
    from jdet_b200.models.roi_heads import OrientedRPNHead
    from jdet_b200.models.boxes import rectpoly2obb
    rpn = OrientedRPNHead(256).to(dev).eval().requires_grad_(False)
    torch.nn.init.normal_(rpn.rpn_cls.weight, 0, 0.05)
    torch.nn.init.normal_(rpn.rpn_reg.weight, 0, 0.02)
    fpn5 = fpn + [torch.randn((2, 256, 16, 16), device=dev, generator=gh)]
    torch.nn.init.normal_(ohead.fc_cls.weight, 0, 0.05)

    send, recv = jdist.gather_buffers(dev, 2, 2000, world)

    def heads_fn():
        props_, counts_ = rpn.forward_batched(fpn5)
        ohead.detect_records(fpn5, props_, counts_, 0.1, 2000, out=send)
        return jdist.all_gather_records(send, recv)

    got = heads_fn()
    K = 10
    ms_heads = agg(time_steps(torch, heads_fn, K, 3, flush)) / K
    ex["oriented_rcnn_heads"] = {
        "metric": "images/s", "value": 2 * world / (ms_heads * 1e-3), "unit": "images/s", "ms_per_step": ms_heads, "steps": K,
        "config": {"workload": "BASELINE configs[4] minus backbone/FPN: 2 x 1024^2 tiles per GPU, OrientedRPNHead over 5 FPN levels "
                               "(nms_pre/post 2000, batched decode, one horizontal NMS call via torchvision) -> OrientedHead (fused "
                               "4-level rotated RoIAlign, 2 shared fp32 FCs, decode, score threshold) -> ONE per-class rotated NMS call "
                               "for both tiles -> records packed into the send buffer -> all-gather of 2 x 2001 x 7 records",
                   "detections_per_image": [int(r[-1, 0].item()) for r in got.reshape(-1, 2001, 7)[:2]]},
        "roofline": {"bound": "hbm", "achieved": 0.0, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": 0.0,
                     "note": "mixed pipeline (cuDNN/cuBLAS convs and FCs around the hot-path ops); no single roofline"}}
    del fpn, fpn5, ohead, rpn

    # cfg5 end to end, caller-true boundary: uint8 tiles up from pinned host memory, detection records down.  The backbone /
    # FPN stand-in (torchvision R50-FPN, random weights) runs in strict fp32 (TF32 off), like the reference's cuDNN path.
    from jdet_b200.models.networks import OrientedRCNN
    prev_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(11 + rank)
        net = OrientedRCNN().to(dev).eval().requires_grad_(False)
        torch.nn.init.normal_(net.rpn.rpn_cls.weight, 0, 0.05); torch.nn.init.normal_(net.rpn.rpn_reg.weight, 0, 0.02)
        torch.nn.init.normal_(net.roi_head.fc_cls.weight, 0, 0.05)
        tiles_h = torch.randint(0, 256, (2, 3, 1024, 1024), dtype=torch.uint8).pin_memory()
        tiles_d = torch.empty_like(tiles_h, device=dev)
        send2, recv2 = jdist.gather_buffers(dev, 2, 2000, world)
        out_h = torch.empty(tuple(recv2.shape) if world > 1 else (1,) + tuple(send2.shape), dtype=torch.float32).pin_memory()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def e2e_fn():
            tiles_d.copy_(tiles_h, non_blocking=True)
            net(tiles_d, out=send2)
            ev[1].record()
            g_ = jdist.all_gather_records(send2, recv2)
            ev[2].record()
            out_h.copy_(g_, non_blocking=True)

        e2e_fn(); torch.cuda.synchronize()
        K = 5
        tot = gat = 0.0
        for it in range(K + 2):
            flush()
            ev[0].record(); e2e_fn(); ev[3].record(); torch.cuda.synchronize()
            if it >= 2:
                tot += ev[0].elapsed_time(ev[3]); gat += ev[1].elapsed_time(ev[2])
        ms_net, ms_gather = agg(tot / K), agg(gat / K)
        ex["oriented_rcnn_e2e"] = {
            "metric": "images/s", "value": 2 * world / (ms_net * 1e-3), "unit": "images/s", "ms_per_step": ms_net, "steps": K,
            "config": {"workload": "BASELINE configs[4]: Oriented R-CNN R50-FPN inference, 2 x 1024^2 uint8 tiles per GPU from pinned "
                                   "host memory (6.3 MB H2D) -> torchvision R50-FPN stand-in (random weights, strict fp32) -> heads as "
                                   "above -> all-gather -> records to the host (%d B D2H)" % (out_h.numel() * 4),
                       "all_gather_ms": ms_gather, "detections_per_image": [int(r[-1, 0].item()) for r in out_h.reshape(-1, 2001, 7)[:2]]},
            "roofline": {"bound": "hbm", "achieved": 0.0, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": 0.0,
                         "note": "whole network; the backbone's cuDNN convolutions dominate (out of the hot-path scope)"}}
        del net, tiles_d
    finally:
        torch.backends.cudnn.allow_tf32 = prev_tf32

    # cfg4: S2ANet-R50-FPN shapes, bs 8: feature_refine + AlignConv over the 5 levels
    levels = [(128, 8), (64, 16), (32, 32), (16, 64), (8, 128)]
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    xs = [torch.randn((8, 256, hw, hw), device=dev, generator=g) for hw, _ in levels]
    an = [cu(s2anet_anchors(rng, 8, hw, hw, st)) for hw, st in levels]
    bx = [a[..., [1, 0, 2, 3, 4]].contiguous() for a in an]
    for points in (1, 5):
        fn = lambda: ops.fr.feature_refine_multi(xs, bx, [1.0 / st for _, st in levels], points)   # the 5 levels in one call
        fn()
        K = 10
        ms = agg(time_steps(torch, fn, K, 3, flush)) / K
        alg = sum(2 * x.numel() * 4 + b.numel() * 4 for x, b in zip(xs, bx))
        ex["feature_refine_p%d" % points] = {
            "metric": "positions/s", "value": sum(x.shape[0] * x.shape[2] * x.shape[3] for x in xs) * world / (ms * 1e-3),
            "unit": "positions/s", "ms_per_step": ms, "steps": K,
            "config": {"workload": "feature_refine points=%d, bs 8 x 256 ch x {128,64,32,16,8}^2 (BASELINE configs[3])" % points},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg}}
    ac = AlignConv(256, 256, 3).to(dev).requires_grad_(False)     # inference: the fused tcgen05 path
    fn = lambda: ac.forward_multi(xs, an, [st for _, st in levels])     # one persistent launch over the 5 levels
    fn()
    K = 5
    ms = agg(time_steps(torch, fn, K, 3, flush)) / K
    flops = 2.0 * 256 * 2304 * sum(x.shape[0] * x.shape[2] * x.shape[3] for x in xs)
    tf32 = measure_tf32_peak(torch, dev)
    ach = flops / (ms * 1e-3) / 1e12
    ex["align_conv"] = {"metric": "positions/s", "value": sum(x.shape[0] * x.shape[2] * x.shape[3] for x in xs) * world / (ms * 1e-3),
                        "unit": "positions/s", "ms_per_step": ms, "steps": K,
                        "config": {"workload": "AlignConv 256->256 3x3, bs 8 x {128,64,32,16,8}^2 (BASELINE configs[3]), the 5 levels in one call"},
                        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf32 / 3.0, "unit": "TFLOP/s", "frac": ach / (tf32 / 3.0),
                                     "flops": flops, "tf32_tflops_measured": tf32, "tensor_tflops_issued": 3.0 * ach,
                                     "note": "achieved = fp32-equivalent FLOPs of the convolution; the kernel issues 3 TF32 MMAs per "
                                             "product (hi*hi + hi*lo + lo*hi, fp32-class accuracy as the reference's SGEMM), so the "
                                             "ceiling is a third of the TF32 rate, measured in this run: torch.matmul 8192^3 with "
                                             "allow_tf32, best of 5"}}
    return ex


def pin_to_gpu_numa_node(index):
    """Bind this process to the CPUs NVML names as local to GPU `index` (its NUMA node), so that pinned host buffers are
    allocated there (first touch) and every rank's PCIe traffic stays on its own socket.  Best effort; returns what it did."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (v >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = None
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
        return {"cpus": len(cpus), "numa_node": node}
    except Exception as e:
        return {"note": "affinity not set: %r" % (e,)}


def measure_tf32_peak(torch, dev):
    """cuBLAS TF32 GEMM rate on this GPU, in this run (the driver's MEASURED_PEAKS.json has no TF32 figure)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn((n, n), device=dev), torch.randn((n, n), device=dev)
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def reference_cuda_legs(torch, dev, flush, roi_ms, extra):
    """the three CUDA-only ops have no reference CPU path — their baseline is the reference's OWN CUDA kernel
    (source strings compiled for sm_100a into oracle/_ref/libref_cuda.so, reference launch geometry) on the same GPU, same
    inputs, same CUDA events and L2 flush.  A reported baseline (rank 0, N = 1); never part of the product path."""
    out = {}
    try:
        import _refcuda
        from _inputs import clustered_boxes, dota_boxes, s2anet_anchors, tie_free_scores
        if not _refcuda.available():
            return {"note": "oracle/_ref/libref_cuda.so did not travel"}
        import ctypes
        import oracle
        R = oracle.ref_cuda()
        rng = np.random.default_rng(100)
        cu = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        pp = lambda t: ctypes.c_void_p(t.data_ptr())
        f32 = lambda v: ctypes.c_float(np.float32(v))

        def leg(name, fn, K, ours_ms, what):
            fn()
            ms = time_steps(torch, fn, K, 2, flush) / K
            out[name] = {"reference_ms": ms, "ours_ms": ours_ms, "ratio": ms / ours_ms, "what": what}

        # RoIAlign cfg2
        g = torch.Generator(device=dev).manual_seed(1234)
        feat = torch.randn((1, 256, 256, 256), device=dev, generator=g)
        rois = cu(make_cfg2(0))
        o = torch.empty((2048, 256, 7, 7), device=dev)
        leg("roi_align_rotated", lambda: R.ref_roi_align_rotated_cuda(1, pp(feat), pp(rois), 2048, 256, 256, 256, 7, 7, f32(0.25), ctypes.c_float(2.0), pp(o), st()),
            20, roi_ms, "ROIAlignRotatedForward v1 (ops/roi_align_rotated_v1.py:70-147), cfg2")
        del feat, o
        # IoU 16k^2 and 1k^2
        for nn, K in ((16384, 3), (1000, 20)):
            b1, b2 = cu(dota_boxes(rng, nn)), cu(dota_boxes(rng, nn))
            o = torch.empty((nn, nn), device=dev)
            key = "box_iou_rotated_%dk" % (nn // 1000)
            leg(key, lambda: R.ref_box_iou_rotated_cuda(pp(b1), nn, pp(b2), nn, pp(o), st()), K, extra[key]["ms_per_step"],
                "box_iou_rotated_cuda_kernel (ops/box_iou_rotated.py:412-461), %dx%d" % (nn, nn))
            del b1, b2, o
        # feature_refine level 0 of cfg4 would need its own ours-leg; use the whole cfg4 (5 levels) like the ours-leg
        levels = [(128, 8), (64, 16), (32, 32), (16, 64), (8, 128)]
        g = torch.Generator(device=dev).manual_seed(77)
        xs = [torch.randn((8, 256, hw, hw), device=dev, generator=g) for hw, _ in levels]
        bx = [cu(s2anet_anchors(rng, 8, hw, hw, s_))[..., [1, 0, 2, 3, 4]].contiguous() for hw, s_ in levels]
        os_ = [torch.empty_like(x) for x in xs]
        for points in (1, 5):
            leg("feature_refine_p%d" % points,
                lambda: [R.ref_feature_refine_cuda(pp(x), pp(b), 8, 256, x.shape[2], x.shape[3], points, f32(1.0 / s_), pp(o_), st())
                         for x, b, o_, (_, s_) in zip(xs, bx, os_, levels)],
                5, extra["feature_refine_p%d" % points]["ms_per_step"], "feature_refine_forward_kernel (ops/fr.py:114-165), cfg4, 5 levels")
        del os_, bx
        # AlignConv: reference = ~25 elementwise offset kernels (not timed) + deformable_im2col + SGEMM per level
        an = [cu(s2anet_anchors(rng, 8, hw, hw, s_)) for hw, s_ in levels]
        from jdet_b200.models.roi_heads.s2anet_head import AlignConv
        acm = AlignConv(256, 256, 3)
        offs = [acm.get_offset_batched(a, s_) for a, (_, s_) in zip(an, levels)]
        w = torch.randn((256, 2304), device=dev) * 0.02
        cols = [torch.empty((2304, 8 * hw * hw), device=dev) for hw, _ in levels]

        def ref_ac():
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            for x, of, col in zip(xs, offs, cols):
                H = x.shape[2]
                R.ref_deformable_im2col_cuda(pp(x), pp(of), 256, H, H, 3, 3, 1, 1, 1, 1, 1, 1, 8, 1, pp(col), st())
                torch.matmul(w, col)
            torch.backends.cuda.matmul.allow_tf32 = prev
        leg("align_conv", ref_ac, 3, extra["align_conv"]["ms_per_step"],
            "deformable_im2col_gpu_kernel (ops/dcn_v1.py:131-184) + fp32 SGEMM (torch.matmul, TF32 off) per level; the reference's "
            "offset micro-kernels are not included")
        del xs, an, offs, cols
        # NMS: the reference mask kernel alone (its host-side greedy pass over a 1.25 GB managed mask is not timed)
        n = 100000
        d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
        s_, l_ = tie_free_scores(rng, n), rng.integers(0, 15, n)
        d6 = cu(np.concatenate([d, l_[:, None].astype(np.float32)], 1))
        order = torch.argsort(cu(s_), descending=True).int()
        keep, ms = _refcuda.nms_rotated_keep(d6, order, 0.1)
        out["nms_rotated"] = {"reference_ms": float(ms), "ours_ms": extra["nms_rotated"]["ms_per_step"],
                              "ratio": float(ms) / extra["nms_rotated"]["ms_per_step"],
                              "what": "nms_rotated_cuda_kernel alone (ops/nms_rotated.py:352-411), 100k x 15, one launch; the "
                                      "reference then reduces a 1.25 GB mask on the host, which is not in this number"}
    except Exception as e:  # a reported baseline, never a reason to fail the bench
        out["note"] = "reference CUDA legs incomplete: %r" % (e,)
    return out


# ------------------------------------------------------------------------------------------------
def cpu_baseline_roi(feat_np, rois_np, budget_s=12.0):
    """Oracle port of ROIAlignRotated_v1 on the host cores (the reference has no CPU path for this op)."""
    import oracle
    sample = 256
    t0 = time.perf_counter()
    oracle.roi_align_rotated(feat_np, rois_np[:sample], (7, 7), 0.25, 2, 1)
    dt = time.perf_counter() - t0
    reps, tot = 0, 0.0
    while tot < budget_s and reps < 20:
        t0 = time.perf_counter()
        oracle.roi_align_rotated(feat_np, rois_np[:sample], (7, 7), 0.25, 2, 1)
        tot += time.perf_counter() - t0
        reps += 1
    per = tot / max(reps, 1) if reps else dt
    out = {"value": sample / per, "unit": "RoIs/s", "cores": oracle.max_threads(), "kind": "port",
           "sample": "%d of the 2048 cfg2 RoIs per repetition, %d repetitions, oracle/oracle.cpp orc_roi_align_rotated (OpenMP)" % (sample, reps)}
    # the two ops that DO have a reference CPU path, timed from the reference's own source when it travelled
    try:
        import ctypes
        from _inputs import clustered_boxes, dota_boxes, tie_free_scores
        RC = oracle.ref_cpu()
        rng = np.random.default_rng(0)
        fp = ctypes.POINTER(ctypes.c_float)
        if RC is not None:
            b1, b2 = dota_boxes(rng, 1000), dota_boxes(rng, 1000)
            o = np.zeros((1000, 1000), np.float32)
            t0 = time.perf_counter()
            RC.ref_box_iou_rotated_cpu(b1.ctypes.data_as(fp), 1000, b2.ctypes.data_as(fp), 1000, o.ctypes.data_as(fp))
            dt = time.perf_counter() - t0
            out["box_iou_rotated_1k"] = {"value": 1e6 / dt, "unit": "pairs/s", "cores": 1, "kind": "reference",
                                         "sample": "cfg1 1000x1000, reference cpu_src (serial, as shipped)"}
            n = 20000
            d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
            d6 = np.concatenate([d, rng.integers(0, 15, n)[:, None].astype(np.float32)], 1).astype(np.float32)
            order = oracle.argsort_desc(tie_free_scores(rng, n))
            keep, sup = np.zeros(n, np.bool_), np.zeros(n, np.uint8)
            t0 = time.perf_counter()
            RC.ref_nms_rotated_cpu(d6.ctypes.data_as(fp), n, 6, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                   ctypes.c_float(0.1), sup.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                   keep.ctypes.data_as(ctypes.POINTER(ctypes.c_bool)))
            dt = time.perf_counter() - t0
            out["nms_rotated_20k"] = {"value": n / dt, "unit": "boxes/s", "cores": 1, "kind": "reference",
                                      "sample": "20k of the cfg3 proposals x 15 classes (O(N^2) serial loop, as shipped)"}
    except Exception as e:  # the baseline is a report, never a reason to fail the bench
        out["note"] = "reference CPU legs skipped: %r" % (e,)
    return out


WORKLOAD = ("roi_align_rotated_v1: 256-ch 256x256 FPN map, 2048 RoIs, 7x7 output, sampling_ratio 2, "
            "spatial_scale 0.25 (BASELINE configs[1]); one tile per GPU")


def run_reference(args):
    """--impl reference: the CPU implementation of the headline path on ALL host cores, the full 2048 RoIs per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)          # torchrun exports OMP_NUM_THREADS=1: this arm must use every core
    import oracle
    rng = np.random.default_rng(1234)
    feat = rng.standard_normal((1, 256, 256, 256)).astype(np.float32)
    rois = make_cfg2(0)
    K, W = min(args.steps, 40), min(args.warmup, 3)
    for _ in range(W):
        oracle.roi_align_rotated(feat, rois, (7, 7), 0.25, 2, 1, threads=ncpu)
    t0 = time.perf_counter()
    for k in range(K):
        oracle.roi_align_rotated(feat, rois, (7, 7), 0.25, 2, 1, threads=ncpu)
    dt = time.perf_counter() - t0
    v = rois.shape[0] * K / dt
    line = {"impl": "reference", "metric": "rotated RoIs/s", "value": v, "unit": "RoIs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rois_per_gpu": int(rois.shape[0])},
            "cpu_baseline": {"value": v, "unit": "RoIs/s", "cores": ncpu, "kind": "port",
                             "sample": "all 2048 RoIs per step x %d steps on %d threads; the reference has no CPU path for "
                                       "roi_align_rotated (CUDA-only), so this is oracle/oracle.cpp (OpenMP over RoIs)" % (K, ncpu)},
            "e2e": {"value": v, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    # Libraries (NCCL prints its version banner to stdout) must not pollute the ONE JSON line the driver
    # parses: route fd 1 to stderr for the whole run and keep the real stdout for the final line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="headline op only (used under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
