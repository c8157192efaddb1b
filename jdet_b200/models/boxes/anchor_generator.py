"""S2ANet grid anchors (reference: python/jdet/models/boxes/anchor_generator.py:113-183)."""
import torch


class AnchorGeneratorRotatedS2ANet:
    def __init__(self, base_size, scales, ratios, angles=[0, ], scale_major=True, ctr=None):
        self.base_size = base_size
        self.scales = torch.tensor(scales, dtype=torch.float32)
        self.ratios = torch.tensor(ratios, dtype=torch.float32)
        self.angles = torch.tensor(angles, dtype=torch.float32)
        self.scale_major = scale_major
        self.ctr = ctr
        self.base_anchors = self.gen_base_anchors()

    @property
    def num_base_anchors(self):
        return self.base_anchors.size(0)

    def gen_base_anchors(self):
        w = h = self.base_size
        if self.ctr is None:
            x_ctr, y_ctr = 0.5 * (w - 1), 0.5 * (h - 1)
        else:
            x_ctr, y_ctr = self.ctr
        h_ratios = torch.sqrt(self.ratios)
        w_ratios = 1 / h_ratios
        assert self.scale_major, "AnchorGeneratorRotated only support scale-major anchors!"
        ones = torch.ones_like(self.angles)
        ws = (w * w_ratios[:, None, None] * self.scales[None, :, None] * ones[None, None, :]).view(-1)
        hs = (h * h_ratios[:, None, None] * self.scales[None, :, None] * ones[None, None, :]).view(-1)
        angles = self.angles.repeat(len(self.scales) * len(self.ratios))
        return torch.stack([x_ctr + torch.zeros_like(ws), y_ctr + torch.zeros_like(ws), ws, hs, angles], dim=-1)

    def grid_anchors(self, featmap_size, stride=16, device=None):
        base = self.base_anchors if device is None else self.base_anchors.to(device)
        feat_h, feat_w = featmap_size
        shift_x = torch.arange(0, feat_w, device=base.device) * stride
        shift_y = torch.arange(0, feat_h, device=base.device) * stride
        xx = shift_x.repeat(len(shift_y))
        yy = shift_y.view(-1, 1).repeat(1, len(shift_x)).view(-1)
        z = torch.zeros_like(xx)
        shifts = torch.stack([xx, yy, z, z, z], dim=-1).to(base.dtype)
        return (base[None, :, :] + shifts[:, None, :]).view(-1, 5)
