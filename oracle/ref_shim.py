"""Import shim that lets the UNMODIFIED reference (prannaykaul/lvc, mounted read-only at
/root/reference) be imported in the build container, where fvcore / iopath / yacs /
pycocotools / termcolor / timm / lvis / matplotlib and the compiled ``detectron2._C`` are absent.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (run in the build container, never on
the GPU box: /root/reference does not travel) to execute the reference's own Python for every row
of SURVEY.md §8(a) and dump golden input/output vectors into ``tests/golden/``.  Nothing in the
product package imports this file.

Only the handful of helpers the model-construction path really needs are implemented for real
(``Registry``, a yaml ``CfgNode`` with ``_BASE_`` inheritance, ``c2_msra_fill``/``c2_xavier_fill``,
``PathManager``); everything else is fabricated as inert placeholder objects.
"""
import copy
import importlib.abc
import importlib.machinery
import os
import sys
import types

import yaml

REFERENCE_ROOT = os.environ.get("LVC_REFERENCE_ROOT", "/root/reference")

_FAKE_ROOTS = (
    "fvcore", "iopath", "yacs", "pycocotools", "termcolor", "timm", "lvis", "imagesize",
    "matplotlib", "mock", "faiss",
)


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


class _Anything(metaclass=_AnyMeta):
    """Placeholder class: subclassable, callable, attribute access returns itself."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _FakeModule(types.ModuleType):
    __all__ = []
    __version__ = "0.1.5"

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Anything,), {})
        setattr(self, name, obj)
        return obj


# ----------------------------------------------------------------------------- real helpers
class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"'{name}' already registered in '{self._name}'"
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(fn):
                self._do_register(fn.__name__, fn)
                return fn
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map


class CfgNode(dict):
    """Minimal yacs/fvcore CfgNode: attribute access, _BASE_ yaml inheritance, merge, freeze."""

    IMMUTABLE = "__immutable__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        d = {}
        for k, v in init_dict.items():
            d[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
        super().__init__(d)
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get(CfgNode.IMMUTABLE, False):
            raise AttributeError(f"Attempted to set {name} on an immutable CfgNode")
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value)
        self[name] = value

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.__dict__[CfgNode.IMMUTABLE]
        out.__dict__[CfgNode.NEW_ALLOWED] = self.__dict__[CfgNode.NEW_ALLOWED]
        return out

    def dump(self, **kwargs):
        def conv(n):
            return {k: conv(v) for k, v in n.items()} if isinstance(n, dict) else n
        return yaml.safe_dump(conv(self), **kwargs)

    @staticmethod
    def load_yaml_with_base(filename, allow_unsafe=False):
        with open(filename, "r") as f:
            cfg = yaml.unsafe_load(f) if allow_unsafe else yaml.safe_load(f)

        def merge_a_into_b(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and k in b:
                    assert isinstance(b[k], dict)
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if "_BASE_" in cfg:
            base = cfg.pop("_BASE_")
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not base.startswith("/"):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base, allow_unsafe=allow_unsafe)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename, allow_unsafe=False):
        loaded = type(self)(CfgNode.load_yaml_with_base(cfg_filename, allow_unsafe))
        self.merge_from_other_cfg(loaded)

    def merge_from_other_cfg(self, other):
        def rec(a, b, path):
            for k, v in a.items():
                if k not in b:
                    if b.__dict__.get(CfgNode.NEW_ALLOWED, False) or True:
                        # the reference's yaml files only set keys its defaults know; be lenient
                        dict.__setitem__(b, k, copy.deepcopy(v))
                        continue
                if isinstance(v, dict) and isinstance(b[k], dict):
                    rec(v, b[k], path + [k])
                else:
                    old = b[k]
                    if isinstance(old, tuple) and isinstance(v, list):
                        v = tuple(v)
                    if isinstance(old, list) and isinstance(v, tuple):
                        v = list(v)
                    if isinstance(old, float) and isinstance(v, int):
                        v = float(v)
                    dict.__setitem__(b, k, copy.deepcopy(v))
        rec(other, self, [])

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            keys = full_key.split(".")
            d = self
            for sub in keys[:-1]:
                d = d[sub]
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except Exception:
                    pass
            old = d.get(keys[-1])
            if isinstance(old, tuple) and isinstance(v, list):
                v = tuple(v)
            if isinstance(old, float) and isinstance(v, int):
                v = float(v)
            dict.__setitem__(d, keys[-1], v)


def _make_weight_init():
    import torch.nn as nn
    m = types.ModuleType("fvcore.nn.weight_init")

    def c2_xavier_fill(module):
        nn.init.kaiming_uniform_(module.weight, a=1)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)

    def c2_msra_fill(module):
        nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)

    m.c2_xavier_fill = c2_xavier_fill
    m.c2_msra_fill = c2_msra_fill
    return m


class _MiniCOCO:
    """The subset of pycocotools.coco.COCO (v2.0) that tools/create_coco_dataset_from_dets_all.py touches: createIndex,
    getImgIds, getAnnIds (imgIds / catIds / areaRng / iscrowd filters), loadAnns.  Semantics follow pycocotools' coco.py."""

    def __init__(self, annotation_file=None):
        self.dataset, self.anns, self.cats, self.imgs = {}, {}, {}, {}
        self.imgToAnns, self.catToImgs = {}, {}
        if annotation_file is not None:
            import json
            self.dataset = json.load(open(annotation_file)) if isinstance(annotation_file, str) else annotation_file
            self.createIndex()

    def createIndex(self):
        from collections import defaultdict
        anns, cats, imgs = {}, {}, {}
        imgToAnns, catToImgs = defaultdict(list), defaultdict(list)
        for ann in self.dataset.get("annotations", []):
            imgToAnns[ann["image_id"]].append(ann)
            anns[ann["id"]] = ann
            catToImgs[ann["category_id"]].append(ann["image_id"])
        for img in self.dataset.get("images", []):
            imgs[img["id"]] = img
        for cat in self.dataset.get("categories", []):
            cats[cat["id"]] = cat
        self.anns, self.imgToAnns, self.catToImgs, self.imgs, self.cats = anns, imgToAnns, catToImgs, imgs, cats

    def getImgIds(self, imgIds=[], catIds=[]):
        return list(self.imgs.keys())

    def getAnnIds(self, imgIds=[], catIds=[], areaRng=[], iscrowd=None):
        imgIds = imgIds if isinstance(imgIds, (list, tuple)) else [imgIds]
        catIds = catIds if isinstance(catIds, (list, tuple)) else [catIds]
        if len(imgIds) == len(catIds) == len(areaRng) == 0:
            anns = self.dataset["annotations"]
        else:
            if len(imgIds) != 0:
                import itertools
                anns = list(itertools.chain.from_iterable(self.imgToAnns[i] for i in imgIds if i in self.imgToAnns))
            else:
                anns = self.dataset["annotations"]
            anns = anns if len(catIds) == 0 else [a for a in anns if a["category_id"] in catIds]
            anns = anns if len(areaRng) == 0 else [a for a in anns if a["area"] > areaRng[0] and a["area"] < areaRng[1]]
        if iscrowd is not None:
            return [a["id"] for a in anns if a["iscrowd"] == iscrowd]
        return [a["id"] for a in anns]

    def loadAnns(self, ids=[]):
        return [self.anns[i] for i in ids] if isinstance(ids, (list, tuple)) else [self.anns[ids]]


class _PathManager:
    @staticmethod
    def open(path, mode="r", **kw):
        return open(path, mode)

    @staticmethod
    def isfile(path):
        return os.path.isfile(path)

    @staticmethod
    def exists(path):
        return os.path.exists(path)

    @staticmethod
    def isdir(path):
        return os.path.isdir(path)

    @staticmethod
    def mkdirs(path):
        os.makedirs(path, exist_ok=True)

    @staticmethod
    def get_local_path(path, **kw):
        return path

    @staticmethod
    def ls(path):
        return os.listdir(path)

    def register_handler(self, *a, **k):
        pass


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _FAKE_ROOTS or fullname == "detectron2._C":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        name = spec.name
        if name == "fvcore.nn.weight_init":
            m = _make_weight_init()
        else:
            m = _FakeModule(name)
        m.__path__ = []
        if name == "fvcore.common.registry":
            m.Registry = Registry
        elif name == "fvcore.common.config":
            m.CfgNode = CfgNode
        elif name in ("fvcore.common.file_io", "iopath.common.file_io"):
            m.PathManager = _PathManager if name.startswith("iopath") else _PathManager()
            m.PathHandler = type("PathHandler", (), {})
            m.HTTPURLHandler = type("HTTPURLHandler", (), {"__init__": lambda s, *a, **k: None})
            m.OneDrivePathHandler = type("OneDrivePathHandler", (), {"__init__": lambda s, *a, **k: None})
        elif name == "fvcore.transforms.transform":
            names = ["BlendTransform", "CropTransform", "PadTransform", "GridSampleTransform",
                     "HFlipTransform", "VFlipTransform", "NoOpTransform", "ScaleTransform",
                     "Transform", "TransformList"]
            for n in names:
                setattr(m, n, type(n, (_Anything,), {}))
            m.__all__ = names
        elif name == "fvcore.nn":
            # fvcore is a third-party dependency that is not vendored in the reference (setup.py: fvcore>=0.1.1).  Its published
            # smooth_l1_loss (fvcore/nn/smooth_l1_loss.py) is restated here so that the reference's own RPN.losses (rpn.py:328-400)
            # can be executed to generate fixtures: beta < 1e-5 -> L1, else 0.5 n^2 / beta below beta and n - 0.5 beta above.
            def smooth_l1_loss(input, target, beta, reduction="none"):
                import torch
                n = torch.abs(input - target)
                loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
                return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
            m.smooth_l1_loss = smooth_l1_loss
        elif name == "termcolor":
            m.colored = lambda s, *a, **k: s
        elif name == "pycocotools.coco":
            m.COCO = _MiniCOCO
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make ``import detectron2`` / ``import lvc`` resolve to the reference tree."""
    global _installed
    if _installed:
        return
    _installed = True
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (only exists in the build container)")
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_ROOT)
    from PIL import Image
    if not hasattr(Image, "LINEAR"):
        Image.LINEAR = Image.BILINEAR


def build_reference_model(config_rel, opts=(), calibrate=True, seed=0):
    """Build a reference model on CPU from a yaml under /root/reference/configs.

    ``calibrate`` applies the SURVEY §8(d) synthetic-weight recipe (every bottleneck's
    conv3.norm.weight = 0.2) so that activations stay finite and RPN yields 1000 proposals.
    """
    install()
    import torch
    from lvc.config import get_cfg
    from lvc.modeling import build_model

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "configs", config_rel))
    cfg.merge_from_list(["MODEL.DEVICE", "cpu"] + list(opts))
    cfg.freeze()
    from lvc.config import set_global_cfg
    set_global_cfg(cfg)
    torch.manual_seed(seed)
    model = build_model(cfg)
    model.eval()
    if calibrate:
        with torch.no_grad():
            for name, buf in model.named_buffers():
                if name.endswith("conv3.norm.weight"):
                    buf.fill_(0.2)
    return cfg, model
