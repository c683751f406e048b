"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product): CPU restatement of the COCO mask API calls the
reference's annotation reader makes -- ``maskUtils.frPyObjects`` / ``merge`` / ``decode`` in
``datasets/reader.py:20-66`` (``read_KINS``, ``read_LVIS``, ``read_COCOA``), the step that produces the N x H x W modal
masks the pairwise-order path consumes.

**Parity unpinned.**  The arithmetic lives in a third-party dependency that is neither vendored under
``/root/reference`` nor installed in this image: ``pycocotools`` (imported at ``datasets/reader.py:5,12``; not listed
in the reference's ``requirements.txt``, so no version is pinned -- 2.0.x at the time of the reference).  This file
restates the published algorithm of its ``common/maskApi.c`` (``rleFrString``, ``rleToString``, ``rleEncode``,
``rleDecode``, ``rleFrPoly``, ``rleMerge`` with ``intersect = 0``) in plain Python / numpy; there is no golden vector
to check it against here, so the tests pin the CUDA / C++ product path to THIS restatement and check the
size-independent properties the format offers (encode -> decode round trips, string round trips, exact areas of
axis-aligned polygons, union = OR of the parts).

Conventions of the format: a mask is stored column-major (Fortran order); ``counts`` alternate runs of 0s and 1s
starting with 0s; the compressed string is a LEB128-like code of the counts, delta-coded against ``counts[i - 2]``
from the fourth count on.
"""
import math

import numpy as np

INT_MIN = -(1 << 31)


def rle_from_string(s):
    """``rleFrString``: compressed ASCII counts -> list of run lengths."""
    if isinstance(s, str):
        s = s.encode("ascii")
    cnts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return cnts


def rle_to_string(cnts):
    """``rleToString`` (used by the tests to make compressed inputs)."""
    out = bytearray()
    for i, c in enumerate(cnts):
        x = int(c)
        if i > 2:
            x -= int(cnts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def rle_encode(mask):
    """``rleEncode``: [h, w] {0,1} -> counts over the column-major pixel order."""
    flat = np.asarray(mask, dtype=np.uint8).T.reshape(-1)   # column-major
    cnts = []
    prev, run = 0, 0
    for v in flat:
        if v != prev:
            cnts.append(run)
            run, prev = 0, v
        run += 1
    cnts.append(run)
    return cnts


def rle_decode(cnts, h, w):
    """``rleDecode``: counts -> [h, w] uint8 {0,1}."""
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in cnts:
        c = int(c)
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape(w, h).T.copy()


def _c_int(x):
    """C's (int) cast of a double on x86-64: truncation toward zero; NaN / out of range -> INT_MIN."""
    if x != x or x >= 2147483648.0 or x <= -2147483649.0:
        return INT_MIN
    return int(x)


def rle_fr_poly(xy, h, w):
    """``rleFrPoly``: polygon [x0, y0, x1, y1, ...] (float) -> counts.  Up-samples by 5, walks every edge with the
    longer axis as the driving variable, keeps the x-crossings, down-samples and turns the sorted crossing list
    into runs."""
    xy = [float(v) for v in xy]
    k = len(xy) // 2
    scale = 5.0
    x = [_c_int(scale * xy[2 * j] + .5) for j in range(k)]
    y = [_c_int(scale * xy[2 * j + 1] + .5) for j in range(k)]
    x.append(x[0])
    y.append(y[0])
    u, v = [], []
    for j in range(k):
        xs, xe, ys, ye = x[j], x[j + 1], y[j], y[j + 1]
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe = xe, xs
            ys, ye = ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx != 0 else float("nan")   # 0 / 0 in C (both deltas zero)
            for d in range(dx + 1):
                t = dx - d if flip else d
                u.append(t + xs)
                v.append(_c_int(ys + s * t + .5))
        else:
            s = (xe - xs) / dy
            for d in range(dy + 1):
                t = dy - d if flip else d
                v.append(t + ys)
                u.append(_c_int(xs + s * t + .5))
    px, py = [], []
    for j in range(1, len(u)):
        if u[j] != u[j - 1]:
            xd = float(u[j] if u[j] < u[j - 1] else u[j] - 1)
            xd = (xd + .5) / scale - .5
            if math.floor(xd) != xd or xd < 0 or xd > w - 1:
                continue
            yd = float(v[j] if v[j] < v[j - 1] else v[j - 1])
            yd = (yd + .5) / scale - .5
            if yd < 0:
                yd = 0.0
            elif yd > h:
                yd = float(h)
            yd = math.ceil(yd)
            px.append(int(xd))
            py.append(int(yd))
    a = sorted([px[j] * h + py[j] for j in range(len(px))] + [h * w])
    prev = 0
    for j in range(len(a)):
        t = a[j]
        a[j] -= prev
        prev = t
    b = [a[0]]
    j = 1
    while j < len(a):
        if a[j] > 0:
            b.append(a[j])
            j += 1
        else:
            j += 1
            if j < len(a):
                b[-1] += a[j]
                j += 1
    return b


def segm_components(segm, h, w):
    """The RLE parts of one ``segmentation`` field, as ``frPyObjects`` builds them: a polygon list gives one RLE per
    polygon (``read_LVIS`` merges them), an uncompressed RLE dict its counts, a compressed one its decoded string."""
    if isinstance(segm, list):
        return [rle_fr_poly(p, h, w) for p in segm]
    counts = segm["counts"]
    if isinstance(counts, (list, tuple, np.ndarray)):
        return [[int(c) for c in counts]]
    return [rle_from_string(counts)]


def decode_segm(segm, h, w):
    """``maskUtils.decode(maskUtils.merge(frPyObjects(segm, h, w)))``: union of the parts, [h, w] uint8."""
    m = np.zeros((h, w), np.uint8)
    for c in segm_components(segm, h, w):
        m |= rle_decode(c, h, w)
    return m
