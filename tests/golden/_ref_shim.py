"""Import the UNMODIFIED reference from /root/reference (build container only; the path does not
exist on the GPU box). Two shims are needed: the package metadata lookup in
yolox/__init__.py:1-2 and the absent pycocotools imported by yolox/data/datasets/coco.py:7."""
import importlib.metadata as md
import sys
import types

REF = "/root/reference"


def import_reference():
    _v = md.version
    md.version = lambda n: "0.4.1" if n == "pixeltable-yolox" else _v(n)
    for name, attrs in {"pycocotools": [], "pycocotools.coco": ["COCO"], "pycocotools.cocoeval": ["COCOeval"],
                        "pycocotools.mask": [], "thop": ["profile"]}.items():
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for a in attrs:
                setattr(mod, a, type(a, (), {}))
            sys.modules[name] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import yolox  # noqa: F401
    from yolox.models import YoloPafpn, YoloxHead, YoloxModule
    from yolox.utils import boxes as ref_boxes

    return dict(YoloPafpn=YoloPafpn, YoloxHead=YoloxHead, YoloxModule=YoloxModule, boxes=ref_boxes)
