"""Shared machinery of the two rotated single-level RoI extractors.

Reference behaviour (python/jdet/models/roi_extractors/{oriented,rbox}_single_level.py): every RoI is
assigned to one FPN level by floor(log2(sqrt(w*h)/finest_scale + 1e-6)) clamped to the level range, the
level's rotated RoIAlign is run on that subset, and results land at the RoI's original row.

Here the subsets are formed by ONE stable sort of the level ids (contiguous slices per level, no boolean
masks), and each level's result is written with a single index_copy_ instead of zeros + masked "+=".
"""
import torch
from torch import nn


def pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class RotatedSingleLevelBase(nn.Module):
    ops_module = None            # jdet_b200.ops.roi_align_rotated[_v1], set by the subclass

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale):
        super().__init__()
        spec = dict(roi_layer)
        kind = spec.pop('type')
        assert hasattr(self.ops_module, kind)
        maker = getattr(self.ops_module, kind)
        self.roi_layers = nn.ModuleList(maker(spatial_scale=1 / s, **spec) for s in featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):     # kept for API compatibility
        spec = dict(layer_cfg)
        maker = getattr(self.ops_module, spec.pop('type'))
        return nn.ModuleList(maker(spatial_scale=1 / s, **spec) for s in featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        """scale < finest: 0; [finest, 2 finest): 1; ... ; clamp to num_levels - 1.  rois: (k,6)."""
        side = (rois[:, 3] * rois[:, 4]).sqrt()
        lvl = torch.log2(side / self.finest_scale + 1e-6).floor()
        return lvl.clamp(0, num_levels - 1).long()

    def _pool_by_level(self, feats, rois, lvls):
        out_hw = pair(self.roi_layers[0].output_size)
        pooled = torch.zeros((rois.shape[0], self.out_channels) + tuple(out_hw), dtype=torch.float32, device=rois.device)
        if rois.shape[0] == 0:
            return pooled
        order = torch.argsort(lvls, stable=True)
        counts = torch.bincount(lvls, minlength=len(feats)).tolist()      # one host sync, like the reference's .any_()
        start = 0
        for level, n in enumerate(counts):
            if n:
                idx = order[start:start + n]
                pooled.index_copy_(0, idx, self.roi_layers[level](feats[level], rois.index_select(0, idx)))
            start += n
        return pooled

    # ---- fused path: one C-ABI call for the whole extractor (jdet_roi_align_rotated_fpn) -----------------------
    def _fused(self, version, feats, rois, ext, rs):
        """None when the fused kernel does not apply (autograd, non-fp32 / non-NCHW-contiguous maps, shapes it refuses);
        the caller then runs the per-level path, which produces the same values."""
        import ctypes
        from ...ops._common import check, lib, scratch, stream_ptr
        feats = list(feats)
        f0 = feats[0]
        if torch.is_grad_enabled() and any(f.requires_grad for f in feats):
            return None
        if not (rois.is_cuda and rois.dtype == torch.float32 and rois.dim() == 2 and rois.shape[1] == 6):
            return None
        if any((not f.is_cuda) or f.dtype != torch.float32 or f.dim() != 4 or not f.is_contiguous() or f.device != f0.device
               or f.shape[:2] != f0.shape[:2] for f in feats):
            return None
        layer = self.roi_layers[0]
        ph, pw = pair(layer.output_size)
        sr = int(layer.sampling_ratio)
        n, (B, C), R = len(feats), f0.shape[:2], rois.shape[0]
        out = torch.empty((R, C, ph, pw), dtype=torch.float32, device=f0.device)
        if R == 0:
            return out
        r = rois.contiguous()
        L = lib()
        ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in feats])
        Hs = (ctypes.c_int * n)(*[f.shape[2] for f in feats])
        Ws = (ctypes.c_int * n)(*[f.shape[3] for f in feats])
        sc = (ctypes.c_float * n)(*[float(l.spatial_scale) for l in self.roi_layers][:n])
        with torch.cuda.device(f0.device):
            ws = scratch(L.jdet_roi_align_rotated_fpn_workspace_bytes(n, B, C, Hs, Ws, R, ph, pw, sr), f0.device)
            rc = L.jdet_roi_align_rotated_fpn(version, ptrs, n, B, C, Hs, Ws, sc, r.data_ptr(), R, ph, pw, sr,
                                              float(ext[0]), float(ext[1]), float(rs[0]), float(rs[1]), float(self.finest_scale),
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(f0.device))
        if rc == -3:                # JDET_ERR_UNSUPPORTED
            return None
        check(rc, "roi_align_rotated_fpn")
        return out

