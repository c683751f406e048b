"""Ground-truth order matrices from InstaOrder's string annotations (host side).

Restates ``InstaOrderDataset.get_gt_ordering`` (reference datasets/reader.py:335-400): the annotation file stores per
image ``occlusion = [{'order': 'i<j'} | {'order': 'i<j & j<i'}]`` and ``depth = [{'order': 'i<j' | 'i=j', 'overlap':
bool, 'count': int}]``; the matrices below are what ``eval_order_recall_precision_f1`` / ``eval_depth_order_whdr`` and
the training label logic consume.  Pure numpy; pinned against the unmodified reference class in tests/test_annotations.py.
"""
import numpy as np


def gt_occlusion_matrix(num, occlusion, rm_bidirec=0):
    """reader.py:340-360.  Quirk kept: with ``rm_bidirec == 1`` a bidirectional entry writes -1 at the indices of the
    PREVIOUS entry (the reference never parses the current one in that branch) and fails when it comes first."""
    m = np.zeros((num, num), dtype=np.int64)
    last = None
    for o in occlusion:
        order = o["order"]
        if "&" in order and rm_bidirec == 1:
            if last is None:
                raise UnboundLocalError("reference reader.py:349 uses idx1 / idx2 before assignment")
            m[last[0], last[1]] = -1
            m[last[1], last[0]] = -1
        elif "&" in order:
            i, j = map(int, order.split(" & ")[0].split("<"))
            m[i, j] = 1
            m[j, i] = 1
            last = (i, j)
        else:
            i, j = map(int, order.split("<"))
            m[i, j] = 1
            last = (i, j)
    return m


def gt_depth_matrices(num, depth, rm_overlap=0):
    """reader.py:362-400: (gt_depth, is_overlap, count), all -1 where there is no annotation; depth[i, j] = 1 and
    depth[j, i] = 0 for 'i<j' (i closer), 2 / 2 for 'i=j'."""
    d = -np.ones((num, num), dtype=np.int64)
    ov = -np.ones((num, num), dtype=np.int64)
    cnt = -np.ones((num, num), dtype=np.int64)
    for e in depth:
        order, is_overlap, count = e["order"], e["overlap"], e["count"]
        ch = "<" if "<" in order else "="
        i, j = map(int, order.split(ch))
        v = -1 if (rm_overlap and is_overlap) else (1 if is_overlap else 0)
        ov[i, j] = ov[j, i] = v
        if ch == "<":
            d[i, j], d[j, i] = 1, 0
        else:
            d[i, j] = d[j, i] = 2
        cnt[i, j] = cnt[j, i] = count
    return d, ov, cnt


def gt_ordering(ann, type, rm_bidirec=0, rm_overlap=0):
    """Same call shape as ``data_reader.get_gt_ordering(imgidx, type, ...)`` but on the image's annotation dict."""
    assert type in ["depth", "occlusion"], "order type should be ond of depth or occlusion"
    num = len(ann["instance_ids"])
    if type == "occlusion":
        return gt_occlusion_matrix(num, ann["occlusion"], rm_bidirec)
    return list(gt_depth_matrices(num, ann["depth"], rm_overlap))
