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


_saved = []   # (module or registry map, name, previous object) for uninstall()


def install(strict: bool = False, architectures: bool = True):
    """Rebind the reference's operator names to lvc_b200.layers and (``architectures``) re-register the mining-path
    meta-architectures -- ``GeneralizedRCNN``, ``GeneralizedRCNNRegOnly``, ``ProposalNetwork`` -- in the reference's
    ``META_ARCH_REGISTRY`` (lvc/modeling/meta_arch/build.py:3-17), so that ``lvc.modeling.build_model(cfg)`` -- and with it
    ``tools/train_net.py --eval-only`` / ``tools/train_net_reg_qe.py --eval-only`` -- constructs the B200 models by name.
    Returns the list of (module, name) pairs patched.  Modules that are not importable in this environment are skipped unless
    ``strict``."""
    from . import layers
    patched = []
    if architectures:
        try:
            build = sys.modules.get("lvc.modeling.meta_arch.build") or importlib.import_module("lvc.modeling.meta_arch.build")
            from .modeling import META_ARCHITECTURES
            omap = build.META_ARCH_REGISTRY._obj_map      # Registry.register() refuses to overwrite; the map is the registry
            for name, cls in META_ARCHITECTURES.items():
                _saved.append((omap, name, omap.get(name)))
                omap[name] = cls
                patched.append(("lvc.modeling.meta_arch.build.META_ARCH_REGISTRY", name))
        except Exception:
            if strict:
                raise
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
                _saved.append((mod, name, getattr(mod, name)))
                setattr(mod, name, impl)
                patched.append((m, name))
    try:   # tools/run_nearest_neighbours.py:142-162, 214-227
        tool = sys.modules.get("tools.run_nearest_neighbours")
        if tool is not None:
            from . import knn
            for name in ("assemble_tensors", "run_nearest_neighbours", "get_nn_class_confirmatory"):
                _saved.append((tool, name, getattr(tool, name)))
                setattr(tool, name, getattr(knn, name))
                patched.append(("tools.run_nearest_neighbours", name))
    except Exception:
        if strict:
            raise
    return patched


def uninstall():
    """Undo install(): put the reference's own objects back (most recent first)."""
    while _saved:
        holder, name, prev = _saved.pop()
        if isinstance(holder, dict):
            if prev is None:
                holder.pop(name, None)
            else:
                holder[name] = prev
        else:
            setattr(holder, name, prev)
