"""Shared host-side plumbing for the op mirrors: tensor checks, stream, scratch."""
import torch

from .. import _lib


def lib():
    return _lib.lib()


def require_cuda(*tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            # mirrors ops/dcn_v1.py:588-589 (NotImplementedError without CUDA); there is no CPU path here
            raise NotImplementedError("jdet_b200 ops are CUDA-only (sm_100a); got a non-CUDA tensor")


def f32c(t):
    """fp32 + contiguous view of t (no copy when it already is)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def scratch(nbytes, device):
    """Scratch from torch's caching allocator: stream-ordered, no cudaMalloc in steady state."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


check = _lib.check
