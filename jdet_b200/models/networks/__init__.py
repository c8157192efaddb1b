from .oriented_rcnn import OrientedRCNN

__all__ = ["OrientedRCNN"]
