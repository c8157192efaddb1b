"""jdet.ops.orn mirror — the two ORN pieces S2ANet's head uses right after AlignConv
(reference: python/jdet/ops/orn.py:595-690; ARF kernel :15-44 / CPU :136-170).

* ``active_rotating_filter(weight, indices)`` — ARF forward: every (out, in, entry) weight is copied to
  nRotation rotated positions given by ``indices`` (1-based).  It permutes ~10^5 weights once per
  forward, so it is plain torch indexing here, not a kernel.
* ``ORConv2d`` — conv2d with the ARF-rotated weight;  ``RotationInvariantPooling`` — max over the
  orientation group (the reference's 1x1 conv there is commented out, :612-614).
* ``rotation_invariant_encoding(input, nOrientation)`` / ``RotationInvariantEncoding`` — RIE (orn.py:291-330, 505-531, 582-593):
  per (batch, feature) the orientation responses are rolled so that the strongest comes first; returns the aligned
  features and the uint8 main directions.  Not on S2ANet's path (it uses the pooling variant); mirrored for completeness.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def active_rotating_filter(weight, indices):
    """weight (nOut, nIn, nOri, kH, kW), indices (nOri, kH, kW, nRot) uint8 1-based ->
    (nOut*nRot, nIn*nOri, kH, kW)   (arf_forward, orn.py:260-269)"""
    assert weight.dim() == 5, "only supports a batch of ARFs."
    n_out, n_in, n_ori, kh, kw = weight.shape
    n_rot = indices.shape[3]
    n_entry = n_ori * kh * kw
    idx = indices.reshape(n_entry, n_rot).to(torch.long) - 1             # [l, k] -> target entry
    w = weight.reshape(n_out, n_in, n_entry)
    out = torch.zeros((n_out, n_rot, n_in, n_entry), dtype=weight.dtype, device=weight.device)
    tgt = idx.t().to(weight.device)[None, :, None, :].expand(n_out, n_rot, n_in, n_entry)
    out.scatter_(3, tgt, w[:, None].expand(n_out, n_rot, n_in, n_entry))  # out[i,k,j,idx[l,k]] = w[i,j,l]
    return out.reshape(n_out * n_rot, n_in * n_ori, kh, kw)


_KERNEL_INDICES = {
    1: {a: (1,) for a in (0, 45, 90, 135, 180, 225, 270, 315)},
    3: {0: (1, 2, 3, 4, 5, 6, 7, 8, 9), 45: (2, 3, 6, 1, 5, 9, 4, 7, 8), 90: (3, 6, 9, 2, 5, 8, 1, 4, 7),
        135: (6, 9, 8, 3, 5, 7, 2, 1, 4), 180: (9, 8, 7, 6, 5, 4, 3, 2, 1), 225: (8, 7, 4, 9, 5, 1, 6, 3, 2),
        270: (7, 4, 1, 8, 5, 2, 9, 6, 3), 315: (4, 1, 2, 7, 5, 3, 8, 9, 6)},
}


def arf_indices(n_ori, n_rot, kernel_size):
    """ORConv2d.get_indices (orn.py:644-678)"""
    kh, kw = kernel_size
    d_ori, d_rot = 360 / n_ori, 360 / n_rot
    ind = torch.zeros((n_ori * kh * kw, n_rot), dtype=torch.uint8)
    for i in range(n_ori):
        for j in range(kh * kw):
            for k in range(n_rot):
                angle = d_rot * k
                layer = (i + math.floor(angle / d_ori)) % n_ori
                ind[i * kh * kw + j, k] = int(layer * kh * kw + _KERNEL_INDICES[kw][angle][j])
    return ind.view(n_ori, kh, kw, n_rot)


class ORConv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size=3, arf_config=None, stride=1, padding=0, dilation=1,
                 groups=1, bias=True):
        self.nOrientation, self.nRotation = _pair(arf_config)
        assert (math.log(self.nOrientation) + 1e-5) % math.log(2) < 1e-3, 'invalid nOrientation {}'.format(self.nOrientation)
        assert (math.log(self.nRotation) + 1e-5) % math.log(2) < 1e-3, 'invalid nRotation {}'.format(self.nRotation)
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.register_buffer("indices", arf_indices(self.nOrientation, self.nRotation, self.kernel_size))
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels, self.nOrientation, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels * self.nRotation))
        n = self.in_channels * self.nOrientation
        for k in self.kernel_size:
            n *= k
        nn.init.normal_(self.weight, 0, math.sqrt(2.0 / n))

    def rotate_arf(self):
        return active_rotating_filter(self.weight, self.indices)

    def forward(self, input):
        return F.conv2d(input, self.rotate_arf(), self.bias, self.stride, self.padding, self.dilation, self.groups)

    execute = forward


def rotation_invariant_encoding(input, nOrientation):
    """input (nBatch, nFeature * nOrientation, 1, 1) -> (aligned, mainDirection (nBatch, nFeature) uint8).
    mainDirection = first index of the maximum over the orientation group (strict '>' scan from -FLT_MAX, orn.py:306-315);
    aligned[.., (l - d) mod nOri] = input[.., l] (:316-325).  Differentiable w.r.t. input (the gather is the
    reference's rie_backward, :332-360)."""
    assert input.dim() == 4, "only supports a batch of RIEs."
    assert input.size(2) == 1 and input.size(3) == 1, "mH x mW should be 1x1."
    n_batch, n_channel = input.size(0), input.size(1)
    n_feature = n_channel // nOrientation
    f = input.reshape(n_batch, n_feature, nOrientation)
    # the reference starts from maxVal = -FLT_MAX and updates on '>': -FLT_MAX / -inf / NaN entries never win; the first maximum does
    key = torch.where(f > -torch.finfo(f.dtype).max, f, torch.full_like(f, -float("inf")))
    d = torch.argmax(key, dim=2)                                  # first occurrence of the maximum
    d = torch.where(torch.isneginf(key.max(dim=2)[0]), torch.zeros_like(d), d)   # no entry above -FLT_MAX: direction stays 0 (zero-initialised output)
    src = (torch.arange(nOrientation, device=f.device)[None, None, :] + d[..., None]) % nOrientation
    aligned = torch.gather(f, 2, src)                             # aligned[a] = f[(a + d) mod nOri]
    return aligned.reshape(input.shape), d.to(torch.uint8)


class RotationInvariantEncoding(nn.Module):
    def __init__(self, nOrientation, return_direction=False):
        super().__init__()
        self.nOrientation = nOrientation
        self.return_direction = return_direction

    def forward(self, input):
        output, d = rotation_invariant_encoding(input, self.nOrientation)
        return (output, d) if self.return_direction else output

    execute = forward


class RotationInvariantPooling(nn.Module):
    def __init__(self, nInputPlane, nOrientation=8):
        super().__init__()
        self.nInputPlane = nInputPlane
        self.nOrientation = nOrientation
        # the reference keeps an unused Conv2d + BatchNorm2d here (orn.py:600-606, "TODO remove this"; its call is commented
        # out, :612-614); the parameters exist in reference checkpoints (`or_pool.conv.*`), so they exist here too
        hidden = int(nInputPlane / nOrientation)
        self.conv = nn.Sequential(nn.Conv2d(hidden, nInputPlane, 1, 1), nn.BatchNorm2d(nInputPlane))

    def forward(self, x):
        N, c, h, w = x.shape
        return x.view(N, -1, self.nOrientation, h, w).max(dim=2)[0]

    execute = forward
