from .s2anet_head import AlignConv, bbox_decode  # noqa: F401
