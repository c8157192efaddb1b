"""CPU-only, world_size 2 over gloo: the sharding + single all-gather logic used at N > 1."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jdet_b200.dist import all_gather_detections, pack_detections, shard_range, unpack_detections


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _make(rank):
    g = torch.Generator().manual_seed(rank)
    n = 50 + 10 * rank
    boxes = torch.rand((n, 5), generator=g)
    scores = torch.rand((n,), generator=g)
    labels = torch.randint(0, 15, (n,), generator=g)
    keep = torch.nonzero(scores > 0.5)[:, 0]
    return boxes, scores, labels, keep


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = all_gather_detections(*_make(rank), max_per_img=40)
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


def test_all_gather_detections_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # every rank holds the same gathered tensor, and block r equals rank r's single-process record
    assert np.array_equal(got[0], got[1])
    for r in range(2):
        want = pack_detections(*_make(r), max_per_img=40).numpy()
        assert np.array_equal(got[0][r], want)
        k = int(want[-1, 0])
        dets = unpack_detections(torch.from_numpy(got[0]))[r]
        assert dets.shape == (k, 7) and bool((dets[:-1, 5] >= dets[1:, 5]).all())


def _worker_records(rank, world, port, q):
    """the fused path's collective: persistent send / recv blocks, two images per rank (records built on the CPU here)"""
    from jdet_b200.dist import all_gather_records, gather_buffers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        send, recv = gather_buffers(torch.device("cpu"), 2, 40, world)
        outs = []
        for step in range(2):                                   # the same buffers serve every step
            for i in range(2):
                send[i].copy_(pack_detections(*_make(10 * step + 2 * rank + i), max_per_img=40))
            outs.append(all_gather_records(send, recv).clone().numpy())
        assert gather_buffers(torch.device("cpu"), 2, 40, world)[0] is send
        q.put((rank, outs))
    finally:
        dist.destroy_process_group()


def test_all_gather_records_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_records, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for step in range(2):
        assert np.array_equal(got[0][step], got[1][step]) and got[0][step].shape == (2, 2, 41, 7)
        for r in range(2):
            for i in range(2):
                want = pack_detections(*_make(10 * step + 2 * r + i), max_per_img=40).numpy()
                assert np.array_equal(got[0][step][r, i], want)
