"""Install liblvcb200's operators INTO an importable reference tree (detectron2 / lvc of prannaykaul/lvc), so that the reference's
own model code -- ROIPooler, RPN post-processing, fast_rcnn_inference, tools/run_nearest_neighbours.py -- calls the B200 kernels
without being edited.  This is the runtime form of the bindings listed in INTEGRATION.md.

    import lvc_b200.compat as compat
    compat.install()          # after `import detectron2`, before building the model

Every name is rebound in each module that imported it by value (``from detectron2.layers import batched_nms`` copies the
reference), which is why the patch walks the known import sites instead of only the defining module.
"""
import importlib
import sys

_SITES = {
    # symbol -> modules of the reference that hold a by-value copy (file:line of the import)
    "batched_nms": ["detectron2.layers.nms", "detectron2.layers", "detectron2.modeling.proposal_generator.proposal_utils",
                    "lvc.modeling.roi_heads.fast_rcnn"],                       # nms.py:10; __init__.py:8; proposal_utils.py:6; fast_rcnn.py:12
    "nms": ["detectron2.layers.nms", "detectron2.layers"],                     # nms.py:7
    "roi_align": ["detectron2.layers.roi_align", "detectron2.layers"],         # roi_align.py:15
    "ROIAlign": ["detectron2.layers.roi_align", "detectron2.layers", "detectron2.modeling.poolers"],   # roi_align.py:63; poolers.py:9
}


def install(strict: bool = False):
    """Rebind the reference's operator names to lvc_b200.layers.  Returns the list of (module, name) pairs patched.
    Modules that are not importable in this environment are skipped unless ``strict``."""
    from . import layers
    patched = []
    for name, mods in _SITES.items():
        impl = getattr(layers, name)
        for m in mods:
            try:
                mod = sys.modules.get(m) or importlib.import_module(m)
            except Exception:
                if strict:
                    raise
                continue
            if hasattr(mod, name):
                setattr(mod, name, impl)
                patched.append((m, name))
    try:   # tools/run_nearest_neighbours.py:142-162, 214-227
        tool = sys.modules.get("tools.run_nearest_neighbours")
        if tool is not None:
            from . import knn
            for name in ("assemble_tensors", "run_nearest_neighbours", "get_nn_class_confirmatory"):
                setattr(tool, name, getattr(knn, name))
                patched.append(("tools.run_nearest_neighbours", name))
    except Exception:
        if strict:
            raise
    return patched
