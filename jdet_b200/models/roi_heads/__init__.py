from .s2anet_head import AlignConv  # noqa: F401
