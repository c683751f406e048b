"""Annotation -> mask producer: the host run-length functions of the C-ABI library against the
CPU restatement of pycocotools' maskApi (oracle/coco_mask_oracle.py, "parity unpinned": pycocotools is not available
here), plus the format's own properties.  The GPU rasteriser is checked in test_gpu_masks.py."""
import ctypes as C

import numpy as np
import pytest

from instaorder_b200 import _lib, masks
from oracle import coco_mask_oracle as M


def random_polygon(rng, h, w, k):
    cx, cy = rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h)
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    r = rng.uniform(0.05, 0.45, k) * min(h, w)
    xy = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], axis=1)
    xy = np.round(xy, 2) + rng.choice([0.0, 0.0, 30.0, -30.0], size=(k, 1))   # some vertices outside the image
    return xy.reshape(-1).tolist()


def test_oracle_properties():
    rng = np.random.RandomState(0)
    for _ in range(20):
        h, w = rng.randint(3, 40), rng.randint(3, 40)
        m = (rng.rand(h, w) > rng.uniform(0.2, 0.8)).astype(np.uint8)
        c = M.rle_encode(m)
        assert sum(c) == h * w
        assert np.array_equal(M.rle_decode(c, h, w), m)
        assert M.rle_from_string(M.rle_to_string(c)) == c
    # axis-aligned squares: half-open pixel intervals, exact areas
    m = M.rle_decode(M.rle_fr_poly([10, 10, 20, 10, 20, 20, 10, 20], 40, 50), 40, 50)
    assert m.sum() == 100 and m[10:20, 10:20].all()
    m = M.rle_decode(M.rle_fr_poly([0, 0, 50, 0, 50, 40, 0, 40], 40, 50), 40, 50)
    assert m.all()


@pytest.mark.parametrize("seed", range(4))
def test_polygon_counts_match_oracle(seed):
    rng = np.random.RandomState(100 + seed)
    for _ in range(60):
        h, w = int(rng.randint(8, 300)), int(rng.randint(8, 300))
        poly = random_polygon(rng, h, w, int(rng.randint(3, 24)))
        if rng.rand() < 0.2:
            poly = poly + poly[:2]            # repeated closing vertex: zero-length edge (0 / 0 slope in the C code)
        got = masks._poly_counts(poly, h, w)
        want = M.rle_fr_poly(poly, h, w)
        assert got.tolist() == want, (h, w, poly)
        assert int(got.astype(np.int64).sum()) == h * w


def test_string_counts_match_oracle():
    rng = np.random.RandomState(7)
    for _ in range(40):
        h, w = int(rng.randint(4, 200)), int(rng.randint(4, 200))
        m = np.zeros((h, w), np.uint8)
        for _ in range(rng.randint(1, 6)):
            y0, x0 = rng.randint(0, h), rng.randint(0, w)
            m[y0:y0 + rng.randint(1, h), x0:x0 + rng.randint(1, w)] = 1
        c = M.rle_encode(m)
        s = M.rle_to_string(c)
        assert masks._string_counts(s, h, w).tolist() == c
        assert masks._string_counts(s.decode("ascii"), h, w).tolist() == c


def test_bad_inputs_fail_loudly():
    out = np.empty(4, np.uint32)
    n = C.c_int(0)
    xy = np.asarray([0, 0, 50, 0, 50, 40, 7, 23, 0, 40], np.float64)
    assert _lib.lib().io_rle_from_polygon(_lib.ptr(xy), 5, 40, 50, _lib.ptr(out), 2, C.byref(n)) < 0   # no room
    assert _lib.lib().io_rle_from_string(b"P", 1, _lib.ptr(out), 4, C.byref(n)) < 0                   # truncated code
    assert masks.mask_to_bbox(np.zeros((4, 5), np.uint8)) == [0, 0, 0, 0]
    m = np.zeros((6, 7), np.uint8); m[2:5, 1:3] = 1
    assert masks.mask_to_bbox(m) == [1, 2, 2, 3]
