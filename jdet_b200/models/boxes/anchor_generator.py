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


class AnchorGenerator:
    """Horizontal multi-level anchors of the Oriented R-CNN RPN (reference: anchor_generator.py:186-420, the
    mmdet-v2 generator): per level, ratios x scales boxes (x1,y1,x2,y2) centred on `center_offset * stride`,
    tiled over the feature grid, anchor index fastest."""

    def __init__(self, strides, ratios, scales, base_sizes=None, scale_major=True, center_offset=0.):
        self.strides = [(s, s) if isinstance(s, (int, float)) else tuple(s) for s in strides]
        self.base_sizes = [min(s) for s in self.strides] if base_sizes is None else list(base_sizes)
        self.ratios = torch.tensor(ratios, dtype=torch.float32)
        self.scales = torch.tensor(scales, dtype=torch.float32)
        self.scale_major, self.center_offset = scale_major, center_offset
        self.base_anchors = [self._base(b) for b in self.base_sizes]

    @property
    def num_levels(self):
        return len(self.strides)

    @property
    def num_base_anchors(self):
        return [b.shape[0] for b in self.base_anchors]

    def _base(self, size):
        hr = torch.sqrt(self.ratios)
        wr = 1 / hr
        if self.scale_major:
            ws, hs = (size * wr[:, None] * self.scales[None]).reshape(-1), (size * hr[:, None] * self.scales[None]).reshape(-1)
        else:
            ws, hs = (size * self.scales[:, None] * wr[None]).reshape(-1), (size * self.scales[:, None] * hr[None]).reshape(-1)
        cx = cy = self.center_offset * size
        return torch.stack([cx - 0.5 * ws, cy - 0.5 * hs, cx + 0.5 * ws, cy + 0.5 * hs], -1)

    def grid_anchors(self, featmap_sizes, device=None):
        out = []
        for (fh, fw), (sw, sh), base in zip(featmap_sizes, self.strides, self.base_anchors):
            base = base.to(device) if device is not None else base
            ys, xs = torch.meshgrid(torch.arange(fh, device=base.device, dtype=torch.float32) * sh,
                                    torch.arange(fw, device=base.device, dtype=torch.float32) * sw, indexing="ij")
            shifts = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4)
            out.append((shifts + base[None]).reshape(-1, 4))
        return out
