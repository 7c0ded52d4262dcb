from .engine import DetectorEngine
from .rcnn import GeneralizedRCNN

__all__ = ["DetectorEngine", "GeneralizedRCNN"]
