from .corrector import BoxCorrectorHead
from .engine import DetectorEngine
from .matcher import Matcher, fast_rcnn_losses, pairwise_iou, rpn_losses
from .postprocessing import detector_postprocess
from .sampling import subsample_labels, subsample_labels_batched, subsample_rpn_labels
from .vit import DinoViT, synthetic_vit_state_dict
from .rcnn import META_ARCHITECTURES, GeneralizedRCNN, GeneralizedRCNNRegOnly, ProposalNetwork

__all__ = ["BoxCorrectorHead", "DetectorEngine", "GeneralizedRCNN", "GeneralizedRCNNRegOnly", "ProposalNetwork", "META_ARCHITECTURES",
           "detector_postprocess", "DinoViT", "synthetic_vit_state_dict", "Matcher", "fast_rcnn_losses", "pairwise_iou", "rpn_losses",
           "subsample_labels", "subsample_labels_batched", "subsample_rpn_labels"]
