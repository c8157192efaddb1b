"""Multi-GPU plumbing for the hot path: image-batch sharding + ONE all-gather of final detections.

The geometry path has no cross-image term (RoIs carry a batch index, NMS is per image, AlignConv /
feature_refine are per pixel), so tiles shard across ranks with no data-path collective; the only
exchange is the gather of each rank's fixed-size padded detection record at the end of a step
(SURVEY.md §8e).  The reference never shards inference (`val/test` are rank-0 only,
runner/runner.py:169-190); the parity criterion is that the gathered detections equal the
single-process run image by image.

Works on any initialised torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the
CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous split of n_items over world ranks; the first (n_items % world) ranks get one more."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(boxes, scores, labels, keep, max_per_img=2000):
    """Fixed-size record (max_per_img + 1, 7): rows [x,y,w,h,theta,score,label] of the kept
    detections in descending score order, zero padded; last row holds the count in column 0."""
    rec = torch.zeros((max_per_img + 1, 7), dtype=torch.float32, device=boxes.device)
    if keep.numel():
        s = scores[keep]
        order = torch.argsort(s, descending=True, stable=True)[:max_per_img]
        idx = keep[order]
        k = idx.numel()
        rec[:k, :5] = boxes[idx]
        rec[:k, 5] = scores[idx]
        rec[:k, 6] = labels[idx].to(torch.float32)
        rec[max_per_img, 0] = float(k)
    return rec


def all_gather_detections(boxes, scores, labels, keep, max_per_img=2000, group=None):
    """One collective per step.  Returns (world, max_per_img + 1, 7); rank r's block is its record."""
    rec = pack_detections(boxes, scores, labels, keep, max_per_img)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rec.unsqueeze(0)
    out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out.view(-1), rec.view(-1), group=group)   # flat views: gloo wants 1-D
    return out


def unpack_detections(gathered):
    """-> list over ranks of (k,7) tensors."""
    res = []
    for r in range(gathered.shape[0]):
        k = int(gathered[r, -1, 0].item())
        res.append(gathered[r, :k])
    return res


# ---- the fused path: NMS epilogue -> persistent send buffer -> one collective ------------------------------------------
_BUFFERS = {}


def gather_buffers(device, images_per_rank, max_per_img, world):
    """Persistent (send, recv) tensors of one rank: send (images_per_rank, max_per_img + 1, 7), recv (world, ...) — allocated
    once per shape, so the NMS epilogue (ops.nms_rotated.ml_nms_rotated_record) writes its records where the collective
    reads them and NCCL sees the same addresses every step."""
    key = (str(device), images_per_rank, max_per_img, world)
    if key not in _BUFFERS:
        send = torch.zeros((images_per_rank, max_per_img + 1, 7), dtype=torch.float32, device=device)
        recv = torch.zeros((world,) + tuple(send.shape), dtype=torch.float32, device=device)
        _BUFFERS[key] = (send, recv)
    return _BUFFERS[key]


def all_gather_records(send, recv, group=None):
    """ONE all_gather_into_tensor of the rank's record block; world 1 (or no process group): a view of `send`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return send.unsqueeze(0)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return recv


def nms_and_gather(per_image, iou_threshold, max_per_img=2000, group=None):
    """per_image: list of (boxes (n,5), scores (n,), labels (n,)) of this rank's images.  Per-class rotated NMS of each
    image written as a record straight into the send buffer, then one collective.
    -> (world, images_per_rank, max_per_img + 1, 7)"""
    from .ops.nms_rotated import ml_nms_rotated_record
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = per_image[0][0].device
    send, recv = gather_buffers(dev, len(per_image), max_per_img, world)
    for i, (b, sc, lb) in enumerate(per_image):
        ml_nms_rotated_record(b, sc, lb, iou_threshold, max_per_img, out=send[i])
    return all_gather_records(send, recv, group)
