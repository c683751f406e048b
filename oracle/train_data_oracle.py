"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's training-side ``__getitem__`` (``patch`` mode) for the
order networks: pair choice, crop geometry with shift / scale augmentation, flip, crop + cv2 resizes, labels and the
random A/B swap.  Only ``tests/`` may import it.

Follows (paths relative to /root/reference), drawing from ``np.random`` in exactly the reference's order:
  * ``datasets/depth_occ_order_dataset.py:142-195`` ``_get_pair``         (twins: depth_order_dataset.py:142-195,
    occ_order_dataset.py:139-191)
  * ``datasets/depth_occ_order_dataset.py:207-252`` ``__getitem__`` (^od), ``depth_order_dataset.py:204-244`` (^d),
    ``occ_order_dataset.py:202-279`` (OrderNet / ^o: 70 % occluding pair, 30 % non-pair)
Pinned by ``oracle/gen_golden_traindata.py`` (the unmodified dataset classes with a mocked annotation reader) ->
``tests/golden/traindata.npz``.
"""
import numpy as np

from oracle import oracle as O


def get_pair(image, modal, bboxes, idx1, idx2, sz, base_aug, phase="train", randshift=True, rng=np.random):
    """_get_pair: returns (modal1 u8 [sz,sz], modal2, rgb fp32 [3,sz,sz], new_bbox, flip)."""
    bbox = O.combine_bbox(np.asarray(bboxes)[(idx1, idx2), :])
    centerx = bbox[0] + bbox[2] / 2.
    centery = bbox[1] + bbox[3] / 2.
    size = max([np.sqrt(bbox[2] * bbox[3] * 2.), bbox[2] * 1.1, bbox[3] * 1.1])
    if phase == "train":
        if randshift:
            centerx += rng.uniform(*base_aug["shift"]) * size
            centery += rng.uniform(*base_aug["shift"]) * size
        size /= rng.uniform(*base_aug["scale"])
    new_bbox = [int(centerx - size / 2.), int(centery - size / 2.), int(size), int(size)]
    m1 = O.resize_nearest(O.crop_padding(modal[idx1], new_bbox, 0), sz, sz)
    m2 = O.resize_nearest(O.crop_padding(modal[idx2], new_bbox, 0), sz, sz)
    flip = bool(base_aug["flip"] and rng.rand() > 0.5)
    rgb = O.resize_cubic_u8(O.crop_padding(image, new_bbox, 0), sz, sz)
    if flip:
        m1, m2, rgb = m1[:, ::-1], m2[:, ::-1], rgb[:, ::-1, :]
    return np.ascontiguousarray(m1), np.ascontiguousarray(m2), O.transform_rgb(np.ascontiguousarray(rgb)), new_bbox, flip


def get_pair_resize(image, modal, idx1, idx2, sz, base_aug, rng=np.random):
    """_get_pair_resize (depth_occ_order_dataset.py:81-99): whole image -> sz x sz, INTER_LINEAR on u8; masks nearest."""
    rgb = O.resize_linear_u8(image, sz, sz)
    m1 = O.resize_nearest(modal[idx1], sz, sz)
    m2 = O.resize_nearest(modal[idx2], sz, sz)
    flip = bool(base_aug["flip"] and rng.rand() > 0.5)
    if flip:
        m1, m2, rgb = m1[:, ::-1], m2[:, ::-1], rgb[:, ::-1, :]
    return np.ascontiguousarray(m1), np.ascontiguousarray(m2), O.transform_rgb(np.ascontiguousarray(rgb)), None, flip


def get_pair_image(image, modal, idx1, idx2, sz, base_aug, rng=np.random):
    """_get_pair_image (depth_occ_order_dataset.py:101-140): centred zero-padded square -> sz x sz."""
    m1 = O.resize_nearest(O.pad_square(modal[idx1]), sz, sz)
    m2 = O.resize_nearest(O.pad_square(modal[idx2]), sz, sz)
    flip = bool(base_aug["flip"] and rng.rand() > 0.5)
    rgb = O.resize_linear_u8(O.pad_square(image), sz, sz)
    if flip:
        m1, m2, rgb = m1[:, ::-1], m2[:, ::-1], rgb[:, ::-1, :]
    return np.ascontiguousarray(m1), np.ascontiguousarray(m2), O.transform_rgb(np.ascontiguousarray(rgb)), None, flip


def get_pair_any(mode, image, modal, bboxes, idx1, idx2, sz, base_aug, rng=np.random):
    if mode == "patch":
        return get_pair(image, modal, bboxes, idx1, idx2, sz, base_aug, rng=rng)
    if mode == "resize":
        return get_pair_resize(image, modal, idx1, idx2, sz, base_aug, rng=rng)
    return get_pair_image(image, modal, idx1, idx2, sz, base_aug, rng=rng)


def depth_label(gt_depth, idx1, idx2):
    """depth_occ_order_dataset.py:229-235: A<B -> 0, A=B -> 2, no annotation -> -1."""
    if gt_depth[idx1, idx2] == -1:
        return -1
    if gt_depth[idx1, idx2] == 1 and gt_depth[idx2, idx1] == 0:
        return 0
    if gt_depth[idx1, idx2] == 2:
        return 2
    raise ValueError("inconsistent depth annotation for (%d, %d)" % (idx1, idx2))


def getitem_od(image, modal, bboxes, idx1, idx2, gt_depth, gt_overlap, gt_count, gt_occ, sz, base_aug, rng=np.random,
               mode="patch"):
    """SupDepthOccOrderDataset.__getitem__ (depth_occ_order_dataset.py:207-252)."""
    m1, m2, rgb, nb, flip = get_pair_any(mode, image, modal, bboxes, idx1, idx2, sz, base_aug, rng=rng)
    lab = depth_label(gt_depth, idx1, idx2)
    count, ovl = gt_count[idx1, idx2], gt_overlap[idx1, idx2]
    a_over_b, b_over_a = gt_occ[idx1, idx2], gt_occ[idx2, idx1]
    if rng.rand() < 0.5:
        return dict(rgb=rgb, modal1=m1, modal2=m2, depth=lab, count=count, overlap=ovl,
                    occ=np.float32([b_over_a, a_over_b]), swapped=False, new_bbox=nb, flip=flip)
    lab = 1 if lab == 0 else lab
    return dict(rgb=rgb, modal1=m2, modal2=m1, depth=lab, count=count, overlap=ovl,
                occ=np.float32([a_over_b, b_over_a]), swapped=True, new_bbox=nb, flip=flip)


def getitem_d(image, modal, bboxes, idx1, idx2, gt_depth, gt_overlap, gt_count, sz, base_aug, rng=np.random):
    """SupDepthOrderDataset.__getitem__ (depth_order_dataset.py:204-244)."""
    m1, m2, rgb, nb, flip = get_pair(image, modal, bboxes, idx1, idx2, sz, base_aug, rng=rng)
    lab = depth_label(gt_depth, idx1, idx2)
    count, ovl = gt_count[idx1, idx2], gt_overlap[idx1, idx2]
    if rng.rand() < 0.5:
        return dict(rgb=rgb, modal1=m1, modal2=m2, depth=lab, count=count, overlap=ovl, swapped=False, new_bbox=nb,
                    flip=flip)
    lab = 1 if lab == 0 else lab
    return dict(rgb=rgb, modal1=m2, modal2=m1, depth=lab, count=count, overlap=ovl, swapped=True, new_bbox=nb, flip=flip)


def getitem_occ(algo, image, modal, bboxes, gt_occ, sz, base_aug, extend_bidirec=False, rng=np.random):
    """SupOcclusionOrderDataset.__getitem__ (occ_order_dataset.py:202-279) after _get_pair_ind (:183-200)."""
    gt_occ = gt_occ.copy()
    np.fill_diagonal(gt_occ, -1)
    pairs, non_pairs = np.where(gt_occ == 1), np.where(gt_occ == 0)
    assert len(pairs[0]) > 0, "the reference re-draws another image when there is no occluding pair"
    label = None
    if rng.rand() < 0.7 or len(non_pairs[0]) == 0:
        r = rng.choice(len(pairs[0]))
        idx1, idx2 = pairs[0][r], pairs[1][r]
        label = 3 if (extend_bidirec and gt_occ[idx2, idx1]) else 1
    else:
        r = rng.choice(len(non_pairs[0]))
        idx1, idx2 = non_pairs[0][r], non_pairs[1][r]
        label = 2
    m1, m2, rgb, nb, flip = get_pair(image, modal, bboxes, idx1, idx2, sz, base_aug, rng=rng)
    a_over_b, b_over_a = gt_occ[idx1, idx2], gt_occ[idx2, idx1]
    swapped = not (rng.rand() < 0.5)
    out = dict(rgb=rgb, modal1=m2 if swapped else m1, modal2=m1 if swapped else m2, swapped=swapped, new_bbox=nb,
               flip=flip, idx=(int(idx1), int(idx2)))
    if algo == "OrderNet":
        out["label"] = (0 if label == 1 else label) if swapped else label
    else:
        out["occ"] = np.float32([a_over_b, b_over_a]) if swapped else np.float32([b_over_a, a_over_b])
    return out
