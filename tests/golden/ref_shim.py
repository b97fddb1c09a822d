"""Import shim for the read-only reference checkout (build container only).

Stubs the plotting / COCO / FLOP-count packages the reference imports at module
level but never touches on the decode path, and restores
``np.int`` which ``core/algorithms/yolo_v8.py:231`` still uses.
Used only by make_golden.py; nothing under tests/ that runs on the GPU box imports it.
"""
import os
import sys
import types

REF = os.environ.get("CVPP_REFERENCE", "/root/reference")


def install():
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference checkout not found at {REF}")
    sys.dont_write_bytecode = True
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # noqa
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    if "matplotlib" not in sys.modules:
        mpl = stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = stub("matplotlib.pyplot")
    if "pycocotools" not in sys.modules:
        pc = stub("pycocotools")
        pc.coco = stub("pycocotools.coco", COCO=object)
        pc.cocoeval = stub("pycocotools.cocoeval", COCOeval=object)
    if "thop" not in sys.modules:
        stub("thop", profile=lambda *a, **k: (0, 0))
    if REF not in sys.path:
        sys.path.insert(0, REF)
