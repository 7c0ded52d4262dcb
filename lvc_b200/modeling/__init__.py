from .corrector import BoxCorrectorHead
from .engine import DetectorEngine
from .matcher import Matcher, fast_rcnn_losses, pairwise_iou, rpn_losses
from .rcnn import GeneralizedRCNN

__all__ = ["BoxCorrectorHead", "DetectorEngine", "GeneralizedRCNN", "Matcher", "fast_rcnn_losses", "pairwise_iou", "rpn_losses"]
