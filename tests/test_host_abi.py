"""CPU-only checks of the C ABI: the library loads, exports every symbol the header declares, and its host-side
geometry / LUT functions reproduce the reference fixtures bit-exactly.  No GPU compute is called here."""
import os
import re

import numpy as np

from instaorder_b200 import _lib, engine
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "instaorder_b200.h")).read()
    return sorted(set(re.findall(r"^IO_API [\w\s\*]+?\b(io_\w+)\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s
    assert sorted(_lib.EXPORTS) == syms           # the ctypes table covers the whole header
    assert L.io_abi_version() == 1


def test_pair_enumerate_and_geometry_match_reference(golden_dir):
    assert [tuple(p) for p in engine.enumerate_pairs(5).tolist()] == O.enumerate_pairs(5)
    assert engine.enumerate_pairs(0).shape == (0, 2) and engine.enumerate_pairs(1).shape == (0, 2)
    z = np.load(os.path.join(golden_dir, "geometry.npz"))
    for b, nb in zip(z["boxes"], z["crops"]):
        assert engine.pair_crop_boxes(b, [[0, 1]])[0].tolist() == list(nb)
    rng = np.random.RandomState(3)
    boxes = np.round(rng.uniform(0, 500, size=(50, 4)), 2)
    assert np.array_equal(engine.expand_bbox(boxes, 3.0), O.expand_bbox(boxes, 3.0))
    ib = rng.randint(1, 400, size=(50, 4))
    assert np.array_equal(engine.expand_bbox(ib, 3.0), O.expand_bbox(ib, 3.0))
    for z_case in ("order_c1_o.npz", "order_c2_od.npz"):
        g = np.load(os.path.join(golden_dir, z_case))
        from oracle import gen_golden
        _, _, bx = gen_golden.build_scene(z_case[6:-4])
        assert np.array_equal(engine.expand_bbox(bx, 3.0), g["boxes_expanded"])


def test_normalize_lut_is_bit_exact():
    mean = np.array(O.DATA_MEAN, dtype=np.float32)
    std = np.array(O.DATA_STD, dtype=np.float32)
    out = np.empty((3, 256), dtype=np.float32)
    _lib.check(_lib.lib().io_normalize_lut(_lib.ptr(mean), _lib.ptr(std), _lib.ptr(out)))
    assert np.array_equal(out, O.normalize_lut())


def test_device_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(RuntimeError):
        engine.OrderEngine([2, 3], 256, 8)
    import ctypes as C
    h = C.c_void_p()
    ncs = np.array([2, 3], dtype=np.int32)
    rc = _lib.lib().io_net_create(_lib.ptr(ncs), 2, 256, 8, C.byref(h))
    assert rc < 0 and "CUDA" in _lib.last_error()


def test_pair_tensor_geometry():
    L = _lib.lib()
    assert L.io_pair_tensor_row_pitch(256) == 264 and L.io_pair_tensor_row_pitch(384) == 392
    assert L.io_pair_tensor_bytes(2, 256) == 2 * 262 * 264 * 16
