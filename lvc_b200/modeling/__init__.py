from .corrector import BoxCorrectorHead
from .engine import DetectorEngine
from .rcnn import GeneralizedRCNN

__all__ = ["BoxCorrectorHead", "DetectorEngine", "GeneralizedRCNN"]
