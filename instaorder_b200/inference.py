"""Drop-in replacements for the pairwise-order entry points of the reference's ``inference.py``.

Same names, argument order and return types as the reference (file:line cited per function); the work is done by
``OrderEngine`` (CUDA kernels behind ``libinstaorder_b200.so``).  ``model`` is one of the wrappers in
``instaorder_b200.models`` (same constructor / ``load_state`` as the reference's ``models.*``).
"""
import collections

import numpy as np

from . import engine as _engine

WHDR_KEYS = ["%s_%s" % (o, e) for o in ("ovlX", "ovlO", "ovlOX") for e in ("eq", "neq", "all")]


def _run(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size):
    if patch_or_image == "orig":       # reference inference.py:401-408: input_size is not used, the image's own size is
        eng = model.engine_for_orig(np.shape(inmodal)[1], np.shape(inmodal)[2])
    else:
        eng = model.engine_for(input_size)
    sc = _engine.Scene(image, inmodal, bboxes)
    return eng.infer_scenes([sc], method, pairs=pairs, patch_or_image=patch_or_image)[0]


def infer_order_sup_occ(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size=256, use_rgb=True):
    """reference inference.py:439-512 -> int [N,N] occlusion order matrix."""
    if method not in ("OrderNet", "InstaOrderNet_o"):
        print("method name should be one of {OrderNet or InstaOrderNet_o}")   # reference :503-505
        return
    if not use_rgb:
        raise NotImplementedError("use_rgb=False (2-channel nets): no shipped config uses it")
    return _run(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size)["occ"]


def infer_order_sup_depth(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size,
                          disp_select_method, use_rgb=True):
    """reference inference.py:515-624 -> (int [N,N] depth order matrix, disp_clipped=None)."""
    if method == "midas_pretrained":       # reference :576-583: always from the disparity map
        if patch_or_image != "resize":
            raise NotImplementedError("midas_pretrained runs in 'resize' mode")
        eng = model.engine_for(input_size, disparity=True)
        order, clipped, _ = eng.disparity_order(_engine.Scene(image, inmodal, bboxes), pairs, disp_select_method)
        return order, clipped
    if method in ("InstaDepthNet_d", "InstaDepthNet_od"):
        if disp_select_method != "":     # reference :589-599: mean / median of the network's disparity inside the masks
            if patch_or_image != "resize":
                raise NotImplementedError("InstaDepthNet runs in 'resize' mode (its shipped config)")
            eng = model.engine_for(input_size, disparity=True)
            order, clipped, _ = eng.disparity_order(_engine.Scene(image, inmodal, bboxes), pairs, disp_select_method)
            return order, clipped
        return _run(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size)["depth"], None
    if method != "InstaOrderNet_d":
        print("method name should be one of {InstaOrderNet_d or midas_pretrained}")   # reference :608-610
        return
    if not use_rgb:
        raise NotImplementedError("use_rgb=False (2-channel nets): no shipped config uses it")
    return _run(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size)["depth"], None


def infer_order_sup_occ_depth(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size,
                              disp_select_method):
    """reference inference.py:349-436 -> (occ_order, depth_order)."""
    if method not in ("InstaOrderNet_od", "InstaDepthNet_od"):
        raise NotImplementedError("%s is outside the pairwise-order hot path (SURVEY.md section 8f)" % method)
    r = _run(model, image, inmodal, bboxes, pairs, method, patch_or_image, input_size)
    return r["occ"], r["depth"]


def infer_gt_order(inmodal, amodal):
    """reference inference.py:719-739 -> int [N,N] ground-truth occlusion order (KINS)."""
    import torch
    from . import _lib
    inmodal = np.ascontiguousarray(inmodal, dtype=np.uint8)
    amodal = np.ascontiguousarray(amodal, dtype=np.uint8)
    n, h, w = inmodal.shape
    pairs = _engine.enumerate_pairs(n)
    dev = torch.device("cuda", torch.cuda.current_device())
    mat = torch.zeros((n, n), dtype=torch.int64, device=dev)
    if pairs.shape[0]:
        m, a, pr = (torch.from_numpy(x).to(dev) for x in (inmodal, amodal, pairs))
        _lib.check(_lib.lib().io_infer_gt_order(m.data_ptr(), a.data_ptr(), n, h, w, pr.data_ptr(), pairs.shape[0],
                                                mat.data_ptr(), _lib.stream_ptr()))
    return mat.cpu().numpy()


def _mask_stats_and_bordering(inmodal, need_bordering):
    """(stats int64 [N, 3] = (sum, #ones, sum of y over ones), bordering bool [P] for the row-major i < j pairs)."""
    import torch
    from . import _lib
    inmodal = np.ascontiguousarray(inmodal, dtype=np.uint8)
    n, h, w = inmodal.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    m = torch.from_numpy(inmodal).to(dev)
    st = torch.empty((n, 3), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().io_mask_stats(m.data_ptr(), n, h, w, st.data_ptr(), _lib.stream_ptr()))
    pairs = _engine.enumerate_pairs(n)
    flags = None
    if need_bordering and pairs.shape[0]:
        pr = torch.from_numpy(pairs).to(dev)
        fl = torch.zeros(pairs.shape[0], dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().io_pair_bordering(m.data_ptr(), n, h, w, pr.data_ptr(), pairs.shape[0], fl.data_ptr(),
                                                _lib.stream_ptr()))
        flags = fl.cpu().numpy().astype(bool)
    return st.cpu().numpy(), pairs, flags


def _heuristic(inmodal, key, first_wins, use_bordering):
    """Shared body of the four baselines: for every pair (optionally only bordering ones) the instance with the
    SMALLER key is ``a``; ``first_wins`` selects order[a, b] = 1, otherwise order[b, a] = 1.  Ties go to (j, i) exactly
    as the reference's ``(i, j) if key_i < key_j else (j, i)``."""
    st, pairs, flags = _mask_stats_and_bordering(inmodal, use_bordering)
    n = st.shape[0]
    order = np.zeros((n, n), dtype=np.int64)
    if key == "area":
        k = st[:, 0].astype(np.float64)
    else:   # mean row index of the pixels == 1 (np.where(mask == 1)[0].mean()): exact int sums -> one fp64 division
        with np.errstate(invalid="ignore", divide="ignore"):
            k = st[:, 2].astype(np.float64) / st[:, 1].astype(np.float64)
    for p, (i, j) in enumerate(pairs):
        if flags is not None and not flags[p]:
            continue
        a, b = (i, j) if k[i] < k[j] else (j, i)
        if first_wins:
            order[a, b] = 1
        else:
            order[b, a] = 1
    return order


def infer_occ_order_area(inmodal, occluder="smaller"):
    """reference inference.py:272-289: among bordering pairs the smaller (or larger) mask occludes."""
    return _heuristic(inmodal, "area", occluder == "smaller", True)


def infer_occ_order_yaxis(inmodal, occluder="lower"):
    """reference inference.py:292-307 (``lower`` = smaller mean y, as the reference names it)."""
    return _heuristic(inmodal, "y", occluder == "lower", True)


def infer_depth_order_area(inmodal, closer="smaller"):
    """reference inference.py:310-328: every pair, the smaller (or larger) mask is closer."""
    return _heuristic(inmodal, "area", closer == "smaller", False)


def infer_depth_order_yaxis(inmodal, closer="lower"):
    """reference inference.py:331-346: ``higher`` = smaller mean y; closer == 'lower' writes order[lower, higher]."""
    return _heuristic(inmodal, "y", closer != "lower", False)


def eval_order_recall_precision_f1(order_matrix, gt_order_matrix, zd):
    """reference inference.py:794-802 -> (recall, precision, f1) x100, python floats."""
    if not np.any(np.asarray(gt_order_matrix) != -1):
        raise ValueError("Found empty input array (no entry with gt != -1)")   # sklearn raises in the reference
    r = _engine.metrics_prf([order_matrix], [gt_order_matrix], zd)[0]
    return float(r[0]), float(r[1]), float(r[2])


def eval_depth_order_whdr(order_matrix, gt_order_ovl_count):
    """reference inference.py:764-791 -> defaultdict(list) with the nine '{ovl}_{eq}' keys."""
    gt, ovl, cnt = gt_order_ovl_count
    r = _engine.metrics_whdr([order_matrix], [gt], [ovl], [cnt])[0]
    out = collections.defaultdict(list)
    for k, v in zip(WHDR_KEYS, r):
        out[k].append(-1 if v == -1.0 else float(v))
    return out
