#!/usr/bin/env python
"""Generate tests/golden/ref_cpu_*.npz from the REFERENCE'S OWN SOURCE (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Runs only in the build container (needs /root/reference
to have been compiled); the committed .npz files are what the tests read everywhere else.

  ref_cpu_iou.npz : boxes1, boxes2, iou_v0_cpu, iou_v1_cpu   (reference cpu_src, std::sort hull)
                    iou_v0_cudavariant, iou_v1_cudavariant     (reference CUDA header run on the host)
  ref_cpu_nms.npz : dets6, scores, order, thr, keep5_cpu, keep6_cpu (reference nms_rotated_cpu, `>=`)
  known_answers.npz: SURVEY.md §8c K1-K3, values produced by the reference CPU source
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402

fp = ctypes.POINTER(ctypes.c_float)
HERE = os.path.dirname(os.path.abspath(__file__))


def gen_boxes(rng, n, extent=192.0, clustered=False):
    if clustered:
        seeds = (n + 7) // 8
        base = np.concatenate([rng.uniform(0, extent, (seeds, 2)),
                               np.exp(rng.uniform(np.log(6), np.log(64), (seeds, 2))),
                               rng.uniform(-np.pi / 2, np.pi / 2, (seeds, 1))], 1)
        b = np.repeat(base, 8, 0)[:n].copy()
        b[:, :2] += rng.normal(0, 0.1, (len(b), 2)) * b[:, 2:4]
        b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (len(b), 2)))
        b[:, 4] += rng.normal(0, 0.1, len(b))
        return b.astype(np.float32)
    return np.concatenate([rng.uniform(0, extent, (n, 2)), np.exp(rng.uniform(np.log(4), np.log(64), (n, 2))),
                           rng.uniform(-np.pi / 2, np.pi / 2, (n, 1))], 1).astype(np.float32)


ADVERSARIAL = np.array([
    [0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0], [0, 0, 2, 2, 0], [1, 0, 2, 2, 0], [2, 0, 2, 2, 0],
    [0, 0, 2, 2, np.pi / 2], [0, 0, 2, 2, np.pi / 4], [0, 0, 1e-3, 5, 0.1], [3, 3, 0, 0, 0],
    [0, 0, 2, 2, 1e-7], [10, 10, 20, 8, 0.3], [12, 9, 18, 10, -0.5], [0, 0, 4, 2, 0], [0, 0, 2, 4, np.pi / 2],
    [100, 100, 50, 1e-4, 0.7], [100, 100, 50, 50, 3.0], [100, 100, 50, 50, -3.0], [1e4, 1e4, 30, 10, 0.2],
    [1e4 + 5, 1e4 - 3, 25, 12, -0.4]], np.float32)


def main():
    assert build_ref.build(), "oracle/_ref could not be built (is /root/reference present?)"
    RC, RG = oracle.ref_cpu(), oracle.ref_cuda()
    rng = np.random.default_rng(20260924)
    b1 = np.concatenate([gen_boxes(rng, 60), gen_boxes(rng, 40, clustered=True), ADVERSARIAL])
    b2 = np.concatenate([gen_boxes(rng, 50), gen_boxes(rng, 32, clustered=True), ADVERSARIAL, b1[:20]])

    def mat(fn):
        out = np.zeros((len(b1), len(b2)), np.float32)
        fn(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2), out.ctypes.data_as(fp))
        return out

    def mat_single(fn):
        out = np.zeros((len(b1), len(b2)), np.float32)
        for i in range(len(b1)):
            for j in range(len(b2)):
                out[i, j] = fn(b1[i].ctypes.data_as(fp), b2[j].ctypes.data_as(fp))
        return out

    np.savez_compressed(os.path.join(HERE, "ref_cpu_iou.npz"), boxes1=b1, boxes2=b2,
                        iou_v0_cpu=mat(RC.ref_box_iou_rotated_cpu), iou_v1_cpu=mat(RC.ref_box_iou_rotated_v1_cpu),
                        iou_v0_cudavariant=mat_single(RG.ref_single_iou_v0_cudavariant_host),
                        iou_v1_cudavariant=mat_single(RG.ref_single_iou_v1_cudavariant_host))

    # NMS through the reference's nms_rotated_cpu source
    n = 600
    d = np.concatenate([gen_boxes(rng, n // 2, 256.0), gen_boxes(rng, n - n // 2, 256.0, clustered=True)])
    labels = rng.integers(0, 4, n).astype(np.float32)
    d6 = np.concatenate([d, labels[:, None]], 1).astype(np.float32)
    scores = rng.permutation(np.linspace(0.05, 1.0, n)).astype(np.float32)
    order = oracle.argsort_desc(scores)
    out = {}
    for thr in (0.1, 0.3, 0.5):
        for bl, dd in ((5, np.ascontiguousarray(d6[:, :5])), (6, d6)):
            keep = np.zeros(n, np.bool_)
            sup = np.zeros(n, np.uint8)
            RC.ref_nms_rotated_cpu(dd.ctypes.data_as(fp), n, bl, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                   ctypes.c_float(np.float32(thr)), sup.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                   keep.ctypes.data_as(ctypes.POINTER(ctypes.c_bool)))
            out["keep%d_cpu_thr%02d" % (bl, int(thr * 10))] = keep
    np.savez_compressed(os.path.join(HERE, "ref_cpu_nms.npz"), dets6=d6, scores=scores, order=order, **out)

    # K1-K3
    k1 = np.array([[0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0]], np.float32)
    k1_out = np.zeros((2, 2), np.float32)
    RC.ref_box_iou_rotated_cpu(k1.ctypes.data_as(fp), 2, k1.ctypes.data_as(fp), 2, k1_out.ctypes.data_as(fp))
    k2 = np.array([[0, 0, 1, 1, 0], [0, 0, .5, .5, .3], [0, 0, .9, .9, 0]], np.float32)
    k2_out = np.zeros((3, 3), np.float32)
    RC.ref_box_iou_rotated_cpu(k2.ctypes.data_as(fp), 3, k2.ctypes.data_as(fp), 3, k2_out.ctypes.data_as(fp))
    k3a, k3b = np.array([10, 10, 20, 8, 0.3], np.float32), np.array([12, 9, 18, 10, -0.5], np.float32)
    k3 = np.float32(RC.ref_single_iou_v0_cpu(k3a.ctypes.data_as(fp), k3b.ctypes.data_as(fp)))
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), k1_boxes=k1, k1_iou=k1_out, k2_dets=k2,
                        k2_scores=np.array([.1, .2, .3], np.float32), k2_labels=np.array([1, 1, 1], np.int32),
                        k2_iou=k2_out, k2_keep=np.array([2]), k3_a=k3a, k3_b=k3b, k3_iou=k3)
    print("K1", k1_out.ravel(), "K2", k2_out.ravel(), "K3", k3)

    # ORN: the reference's ARF / RIE CPU sources (ops/orn.py:136-170, 291-330) on small seeded inputs
    import torch
    from oracle import glue
    from jdet_b200.ops.orn import arf_indices
    rng = np.random.default_rng(5)
    orn = {}
    for t, (no, ni, nori, nrot, k) in enumerate(((4, 3, 1, 8, 3), (2, 3, 8, 8, 3), (3, 2, 4, 4, 3), (2, 2, 8, 8, 1))):
        w = rng.standard_normal((no, ni, nori, k, k)).astype(np.float32)
        ind = arf_indices(nori, nrot, (k, k)).numpy()
        orn["arf%d_w" % t], orn["arf%d_ind" % t], orn["arf%d_out" % t] = w, ind, glue.ref_arf_forward(w, ind)
    for t, (nb, nf, nori) in enumerate(((5, 7, 8), (3, 4, 4), (2, 16, 8))):
        f = rng.standard_normal((nb, nf * nori)).astype(np.float32)
        f[0, :nori] = f[0, 0]                                   # a tie: the first maximum wins
        d, al = glue.ref_rie_forward(f, nori)
        orn["rie%d_f" % t], orn["rie%d_nori" % t], orn["rie%d_dir" % t], orn["rie%d_out" % t] = f, np.int32(nori), d, al
    np.savez_compressed(os.path.join(HERE, "ref_cpu_orn.npz"), **orn)
    print("ORN golden:", sorted(orn)[:4], "...")


if __name__ == "__main__":
    main()
