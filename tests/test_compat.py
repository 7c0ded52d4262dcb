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
    compat.uninstall()
    assert pu.batched_nms is not layers.batched_nms


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_build_model_returns_b200_architectures_and_state_dicts_interchange():
    """Registry-level drop-in (lvc/modeling/meta_arch/build.py:3-17): after compat.install() the reference's own build_model(cfg)
    constructs the lvc_b200 nn.Modules by name, and state dicts move between the reference's modules and ours with strict=True in
    both directions (same parameter / buffer names and shapes).  The forward itself needs a GPU, which the
    build container does not have: numerical parity of these models is covered by the -m gpu tests against reference-generated fixtures."""
    import torch
    from oracle import ref_shim
    from lvc_b200 import compat
    from lvc_b200.modeling import GeneralizedRCNN, GeneralizedRCNNRegOnly, ProposalNetwork
    cfg, ref = ref_shim.build_reference_model("COCO-detection/faster_rcnn_R_50_FPN_ft_all_30shot_aug_ftmore_dropout.yaml", calibrate=False)
    compat.install()
    try:
        from lvc.modeling import build_model
        ours = build_model(cfg)
        assert isinstance(ours, GeneralizedRCNN) and isinstance(ours, torch.nn.Module) and not ours.training
        assert ours.device == torch.device("cpu") and ours.cfg.output_layer == "CosineSimOutputLayers"
        ref_sd = ref.state_dict()
        assert set(ref_sd) == set(ours.state_dict()), set(ref_sd) ^ set(ours.state_dict())
        assert all(ref_sd[k].shape == v.shape for k, v in ours.state_dict().items())
        ours.load_state_dict(ref_sd, strict=True)                       # reference checkpoint -> B200 model
        ref.load_state_dict(ours.state_dict(), strict=True)             # and back
        assert torch.equal(ours.state_dict()["roi_heads.box_head.fc1.weight"], ref_sd["roi_heads.box_head.fc1.weight"])
        ours.eval()
        with pytest.raises(NotImplementedError):
            ours.train()
        for arch, cls in (("ProposalNetwork", ProposalNetwork),):
            c2 = cfg.clone()
            c2.defrost()
            c2.MODEL.META_ARCHITECTURE = arch
            m = build_model(c2)
            assert isinstance(m, cls)
        # the box corrector's config (cascade heads, 3 FCs)
        cfg3, ref3 = ref_shim.build_reference_model("COCO-detection/cascade_ubbr_R_50_FPN_ft_all_30shot_aug_ftmore.yaml",
                                                    ["QUERY_EXPAND.ENABLED", True, "MODEL.META_ARCHITECTURE", "GeneralizedRCNNRegOnly"],
                                                    calibrate=False)
        m3 = build_model(cfg3)
        assert isinstance(m3, GeneralizedRCNNRegOnly)
        r3 = ref3.state_dict()
        assert set(r3) == set(m3.state_dict()), sorted(set(r3) ^ set(m3.state_dict()))[:10]
        m3.load_state_dict(r3, strict=True)
    finally:
        compat.uninstall()


def test_layer_wrapper_mirrors_host_side():
    """lvc_b200.layers mirrors of detectron2/layers (shape_spec.py, wrappers.py:14-118, batch_norm.py:13-135): construction, state-dict
    names / versions, helper semantics, and loud refusal of what the GEMM path does not cover (no silent fallback)."""
    import pytest
    import torch
    from lvc_b200 import _lib
    from lvc_b200.layers import Conv2d, FrozenBatchNorm2d, Linear, ShapeSpec, cat, get_norm, nonzero_tuple
    s = ShapeSpec(channels=256, stride=4)
    assert s.channels == 256 and s.height is None and s._replace(stride=8).stride == 8
    a = torch.arange(6).view(2, 3)
    assert cat([a]) is a and cat([a, a], dim=1).shape == (2, 6)
    assert [t.tolist() for t in nonzero_tuple(a > 2)] == [[1, 1, 1], [0, 1, 2]] and nonzero_tuple(torch.tensor(5))[0].tolist() == [0]
    conv = Conv2d(64, 128, kernel_size=3, stride=1, padding=1, bias=False, norm=get_norm("FrozenBN", 128), activation=torch.nn.functional.relu)
    assert sorted(conv.state_dict()) == ["norm.bias", "norm.running_mean", "norm.running_var", "norm.weight", "weight"]
    assert FrozenBatchNorm2d._version == 3 and get_norm("", 8) is None
    bn = FrozenBatchNorm2d(4)
    old = {"weight": torch.ones(4), "bias": torch.zeros(4), "running_mean": torch.zeros(4), "running_var": torch.ones(4)}
    meta = torch.nn.modules.module.OrderedDict()
    old = torch.nn.modules.module.OrderedDict(old)
    old._metadata = {"": {"version": 2}}
    bn.load_state_dict(old)                                    # version 2 stored running_var + eps (batch_norm.py:78-84)
    assert torch.allclose(bn.running_var, torch.ones(4) - 1e-5)
    x = torch.randn(2, 4, 5, 5)
    sc, sh = bn.scale_shift()
    assert torch.allclose(bn(x), x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    with pytest.raises(_lib.LvcB200Error):
        get_norm("BN", 8)
    with pytest.raises(_lib.LvcB200Error):
        Conv2d(3, 64, kernel_size=7, stride=2, padding=3)(torch.zeros(1, 3, 8, 8))      # unsupported shape: refused, not emulated
    with pytest.raises(_lib.LvcB200Error):
        conv(torch.zeros(1, 64, 8, 8))                                                  # CPU tensor: no CPU path
    assert sorted(Linear(16, 8).state_dict()) == ["bias", "weight"]
