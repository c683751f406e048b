"""Annotation -> modal masks on the GPU.

Mirrors the mask-producing calls of the reference's annotation reader -- ``read_KINS`` / ``read_LVIS`` / ``read_COCOA``
(``datasets/reader.py:20-66``): ``maskUtils.frPyObjects`` + ``merge`` + ``decode`` -- with the same argument meaning and
return values, but the N x H x W tensor is produced by ``io_masks_from_rle`` directly in HBM from the run lengths (a
few KB per instance cross PCIe instead of H*W bytes; ``engine.Scene`` accepts the CUDA tensor as is).  Run lengths of
polygons / compressed strings are computed on the host inside the C-ABI library (``io_rle_from_polygon``,
``io_rle_from_string``).  No CPU fallback: without the library every call raises."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _poly_counts(poly, h, w):
    xy = np.ascontiguousarray(np.asarray(poly, dtype=np.float64).reshape(-1))
    k = xy.size // 2
    # every crossing gives at most one run boundary: 5 * perimeter is a generous bound, grown on demand
    cap = 4096
    while True:
        out = np.empty(cap, dtype=np.uint32)
        n = C.c_int(0)
        rc = _lib.lib().io_rle_from_polygon(_lib.ptr(xy), k, int(h), int(w), _lib.ptr(out), cap, C.byref(n))
        if rc == 0:
            return out[:n.value].copy()
        if cap >= (1 << 24):
            _lib.check(rc)
        cap *= 8


def _string_counts(s, h, w):
    if isinstance(s, str):
        s = s.encode("ascii")
    out = np.empty(len(s) + 1, dtype=np.uint32)
    n = C.c_int(0)
    _lib.check(_lib.lib().io_rle_from_string(s, len(s), _lib.ptr(out), out.size, C.byref(n)))
    return out[:n.value].copy()


def segm_components(segm, h, w):
    """Run lengths of the RLE parts of one ``segmentation`` field (polygon list -> one part per polygon, as
    ``frPyObjects`` does; uncompressed / compressed RLE dict -> one part)."""
    if isinstance(segm, list):
        if len(segm) and not isinstance(segm[0], (list, tuple, np.ndarray)):
            segm = [segm]      # a single flat polygon
        return [_poly_counts(p, h, w) for p in segm]
    counts = segm["counts"]
    if isinstance(counts, (list, tuple, np.ndarray)):
        return [np.asarray(counts, dtype=np.uint32)]
    return [_string_counts(counts, h, w)]


def rasterize(segms, h, w, device=None, stream=None):
    """``np.array([decode(merge(frPyObjects(s, h, w))) for s in segms])`` as a CUDA uint8 tensor [N, h, w]."""
    parts, comp_off, inst_off = [], [0], [0]
    for s in segms:
        for c in segm_components(s, h, w):
            tot = int(c.astype(np.int64).sum())
            if tot != h * w:
                raise ValueError("RLE covers %d pixels, image has %d x %d" % (tot, h, w))
            parts.append(np.cumsum(c.astype(np.int64)).astype(np.uint32))
            comp_off.append(comp_off[-1] + c.size)
        inst_off.append(len(comp_off) - 1)
    n = len(segms)
    if device is None:       # the process's current GPU (cuda:LOCAL_RANK once an engine exists), never a fixed cuda:0
        device = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    if n == 0:
        return out
    cum = np.concatenate(parts) if parts else np.zeros(1, np.uint32)
    d_cum = torch.from_numpy(cum.view(np.int32)).to(device)
    d_comp = torch.from_numpy(np.asarray(comp_off, dtype=np.int32)).to(device)
    d_inst = torch.from_numpy(np.asarray(inst_off, dtype=np.int32)).to(device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().io_masks_from_rle(d_cum.data_ptr(), d_comp.data_ptr(), d_inst.data_ptr(), n, int(h), int(w),
                                                out.data_ptr(), _lib.stream_ptr(stream)))
    return out


def decode(segm, h=None, w=None, device=None):
    """``maskUtils.decode`` of one segmentation (RLE dicts carry their own ``size``) -> numpy [h, w] uint8."""
    if isinstance(segm, dict) and "size" in segm:
        h, w = int(segm["size"][0]), int(segm["size"][1])
    return rasterize([segm], h, w, device)[0].cpu().numpy()


# ---- the reader's per-annotation functions (same names, arguments and return values) ---------------------------
def read_KINS(ann, device=None):
    """reference datasets/reader.py:20-28"""
    modal = decode(ann["inmodal_seg"], device=device)
    score = ann["score"] if "score" in ann.keys() else 1.
    return modal, ann["inmodal_bbox"], ann["category_id"], score


def read_LVIS(ann, h, w, device=None):
    """reference datasets/reader.py:31-46"""
    return decode(ann["segmentation"], h, w, device), ann["bbox"], ann["category_id"]


def mask_to_bbox(mask):
    """xywh box of the pixels equal to 1, [0, 0, 0, 0] for an empty mask (reference utils/data_utils.py:75-84)."""
    ys, xs = np.nonzero(np.asarray(mask) == 1)
    if ys.size == 0:
        return [0, 0, 0, 0]
    return [int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)]


def read_COCOA(ann, h, w, device=None):
    """reference datasets/reader.py:49-66"""
    if "visible_mask" in ann.keys():
        modal = decode(ann["visible_mask"], h, w, device)
    else:
        modal = decode([ann["segmentation"]], h, w, device)
    if np.all(modal != 1):
        bbox = mask_to_bbox(decode([ann["segmentation"]], h, w, device))
    else:
        bbox = mask_to_bbox(modal)
    return modal, bbox, 1


def image_instances(anns, h, w, device=None):
    """The mask / box / category triple of ``InstaOrderDataset.get_image_instances`` (reference
    datasets/reader.py:421-457) for the annotations of one image, masks left on the GPU: (uint8 CUDA [N, h, w],
    float64 [N, 4], int64 [N])."""
    masks = rasterize([a["segmentation"] for a in anns], h, w, device)
    boxes = np.asarray([a["bbox"] for a in anns], dtype=np.float64).reshape(-1, 4)
    cats = np.asarray([a["category_id"] for a in anns], dtype=np.int64)
    return masks, boxes, cats
