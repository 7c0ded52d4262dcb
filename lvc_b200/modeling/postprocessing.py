"""detector_postprocess (detectron2/modeling/postprocessing.py:10-79) for the results that do not come out of the fused
``det_merge`` kernel (proposals of ``ProposalNetwork``, corrected boxes of ``GeneralizedRCNNRegOnly``): scale the boxes from
the network input size to the requested output size, clip, drop empty boxes.  At most a few hundred boxes per image."""
from ..structures import Instances


def detector_postprocess(results: Instances, output_height, output_width) -> Instances:
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances((output_height, output_width), **results.get_fields())
    output_boxes = None
    for field in ("pred_boxes", "proposal_boxes", "gt_boxes"):   # the reference rescales every box field it finds, filters on the last
        if results.has(field):
            output_boxes = results.get(field).clone()
            output_boxes.scale(scale_x, scale_y)
            output_boxes.clip(results.image_size)
            results.set(field, output_boxes)
    if output_boxes is None:
        return results
    return results[output_boxes.nonempty()]
