"""compat.install() rebinds the reference's operator names (only checkable where the reference tree exists: the build container)."""
import os
import sys

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_install_patches_reference_import_sites():
    from oracle import ref_shim
    ref_shim.install()
    import detectron2.layers  # noqa: F401
    import detectron2.modeling.poolers  # noqa: F401
    import detectron2.modeling.proposal_generator.proposal_utils as pu
    import lvc.modeling.roi_heads.fast_rcnn as fr
    from lvc_b200 import compat, layers
    patched = compat.install()
    assert ("detectron2.layers.nms", "batched_nms") in patched
    assert pu.batched_nms is layers.batched_nms and fr.batched_nms is layers.batched_nms
    assert sys.modules["detectron2.modeling.poolers"].ROIAlign is layers.ROIAlign
    assert sys.modules["detectron2.layers"].roi_align is layers.roi_align
