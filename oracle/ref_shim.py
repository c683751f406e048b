"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference at /root/reference.

The reference (POSTECH-CVLab/InstaOrder) is pure Python written against torch 1.7 / numpy<1.20 and it
calls ``.cuda()`` unconditionally.  This module makes it importable on a CPU-only box without editing it
(SURVEY.md section 8c lists the shims).  It is used by ``oracle/gen_golden.py`` to produce the committed
fixtures under ``tests/golden/`` and by the ``not gpu`` tests that pin ``oracle/oracle.py`` against the
real reference when ``/root/reference`` exists (build container only -- the GPU box does not have it).

Nothing under ``instaorder_b200/`` may import this file.
"""
import os
import sys
import types

REFERENCE_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "instaorder_ref.zip")


def _root():
    """The unmodified tree when it exists (build container), else the archive of the same modules that
    oracle/build_ref.py packed (GPU box: ``oracle/_ref/`` travels with the snapshot, ``/root/reference`` does not)."""
    r = os.environ.get("INSTAORDER_REFERENCE", "/root/reference")
    if os.path.isfile(os.path.join(r, "inference.py")):
        return r
    if os.path.isfile(REFERENCE_ZIP):
        return REFERENCE_ZIP
    return r


REFERENCE_ROOT = _root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "inference.py")) or REFERENCE_ROOT == REFERENCE_ZIP


def is_archive() -> bool:
    return REFERENCE_ROOT == REFERENCE_ZIP


_loaded = {}


class cpu_only(object):
    """Context manager: every ``.cuda()`` of the reference (it calls them unconditionally: models/single_stage_model.py:26,
    models/supervised_order.py:34-48, utils/data_utils.py:34, inference.py:141-142, utils/common_utils.py:129-130) is a
    no-op inside, so the UNMODIFIED reference runs on the host cores even on a box that has a GPU (bench.py's CPU arm).
    The patches are undone on exit: this package's own ``.cuda()`` / ``.to(device)`` calls are not affected outside."""

    def __enter__(self):
        import torch
        import torch.nn as nn
        self._saved = (torch.Tensor.cuda, nn.Module.cuda, torch.UntypedStorage.cuda, torch.TypedStorage.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        torch.UntypedStorage.cuda = lambda self, *a, **k: self
        torch.TypedStorage.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        import torch
        import torch.nn as nn
        torch.Tensor.cuda, nn.Module.cuda, torch.UntypedStorage.cuda, torch.TypedStorage.cuda = self._saved
        return False


class _Stub(types.ModuleType):
    """Module whose every attribute is a harmless callable (plotting / dataset-only imports)."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return lambda *a, **k: None


def _stub(name, attrs=()):
    sys.modules.setdefault(name, _Stub(name))
    return sys.modules[name]


def load():
    """Returns a namespace with the reference's ``utils``, ``inference`` and ``models`` modules."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np
    import torch
    import torch.nn as nn
    import cv2

    # The open-source OpenCV resize is the pinned arithmetic; the bundled IPP HAL gives results that
    # differ by +-1 u8 LSB on ~3-5 % of INTER_CUBIC pixels and may change with the host CPU.
    cv2.ipp.setUseIPP(False)

    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; reference uses np.int as a dtype
    # stubs for plotting / dataset-only dependencies that are not on the hot path
    _stub("matplotlib")
    sys.modules["matplotlib"].pyplot = _stub("matplotlib.pyplot")
    sk = _stub("skimage")
    sk.io = _stub("skimage.io")
    sk.draw = _stub("skimage.draw")
    sk.morphology = _stub("skimage.morphology", ["convex_hull"])
    pc = _stub("pycocotools")
    pc.mask = _stub("pycocotools.mask")
    pc.coco = _stub("pycocotools.coco", ["COCO"])
    pc.cocoeval = _stub("pycocotools.cocoeval", ["COCOeval"])
    _stub("cvbase")
    # .cuda() is a no-op on the CPU oracle
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        torch.UntypedStorage.cuda = lambda self, *a, **k: self      # utils/common_utils.py:129-130 map_location
        torch.TypedStorage.cuda = lambda self, *a, **k: self

    sys.path.insert(0, REFERENCE_ROOT)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("utils", "inference", "models", "datasets", "midas") or
             k.startswith(("utils.", "models.", "datasets.", "midas."))}
    try:
        import utils as r_utils  # noqa: E402  (import order matters: utils, inference, models)
        import inference as r_inference
        import models as r_models
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(utils=r_utils, inference=r_inference, models=r_models)
    # keep the reference modules reachable only through ``ns`` so that our own package can own these
    # top-level names in the same process
    for k in list(sys.modules):
        if k in ("utils", "inference", "models", "datasets", "midas") or \
                k.startswith(("utils.", "models.", "datasets.", "midas.")):
            sys.modules["_instaorder_ref." + k] = sys.modules.pop(k)
    sys.modules.update(saved)
    _loaded["ns"] = ns
    return ns
