"""Host logic of ``instaorder_b200.reader.InstaOrderDataset`` (file lookup rule, indices, GT matrices, instance
assembly) on synthetic annotation files.  The GPU rasteriser is replaced by the CPU mask oracle here (the kernel itself
is checked in tests/test_gpu_masks.py)."""
import json

import numpy as np
import torch

from instaorder_b200 import masks, reader
from oracle import coco_mask_oracle as M


def write_files(tmp_path):
    h, w = 60, 80
    polys = [[[5, 5, 40, 8, 30, 45]], [[20, 10, 70, 12, 60, 50, 25, 40]], [[50, 30, 75, 30, 75, 55, 50, 55]]]
    rle = dict(size=[h, w], counts=M.rle_to_string(M.rle_encode(M.decode_segm(polys[2], h, w))).decode("ascii"))
    coco = dict(images=[dict(id=7, file_name="000000000007.jpg", width=w, height=h)],
                annotations=[dict(id=100, image_id=7, segmentation=polys[0], bbox=[5, 5, 35, 40], category_id=3),
                             dict(id=101, image_id=7, segmentation=polys[1], bbox=[20, 10, 50, 40], category_id=1),
                             dict(id=102, image_id=7, segmentation=rle, bbox=[50, 30, 25, 25], category_id=9)])
    inst = dict(annotations=[dict(image_id=7, instance_ids=["100", "101", "102"],
                                  occlusion=[dict(order="0<1"), dict(order="1<2 & 2<1")],
                                  depth=[dict(order="0<1", overlap=True, count=2), dict(order="1=2", overlap=False, count=1)])])
    (tmp_path / "instances_val2017.json").write_text(json.dumps(coco))
    p = tmp_path / "InstaOrder_val2017.json"
    p.write_text(json.dumps(inst))
    return str(p), polys, rle, (h, w)


def test_reader_interface(tmp_path, monkeypatch):
    annot_fn, polys, rle, (h, w) = write_files(tmp_path)
    monkeypatch.setattr(masks, "rasterize", lambda segms, hh, ww, device="cpu", stream=None:
                        torch.from_numpy(np.stack([M.decode_segm(s, hh, ww) for s in segms])))
    ds = reader.InstaOrderDataset(annot_fn, device="cpu")
    assert len(ds) == ds.get_image_length() == 1
    modal, category, bboxes, amodal, image_fn = ds.get_image_instances(0, with_gt=True)
    assert image_fn == "000000000007.jpg" and modal.shape == (3, h, w) and modal.dtype == np.uint8
    assert np.array_equal(modal[0], M.decode_segm(polys[0], h, w)) and np.array_equal(modal[2], M.decode_segm(rle, h, w))
    assert category.tolist() == [3, 1, 9] and bboxes.shape == (3, 4) and amodal.size == 0
    assert ds.get_image_instances(0, with_id=True)[-1] == 7
    occ = ds.get_gt_ordering(0, "occlusion")
    assert occ[0, 1] == 1 and occ[1, 0] == 0 and occ[1, 2] == 1 and occ[2, 1] == 1
    depth, ovl, cnt = ds.get_gt_ordering(0, "depth")
    assert depth[0, 1] == 1 and depth[1, 0] == 0 and depth[1, 2] == 2 and ovl[0, 1] == 1 and cnt[0, 1] == 2 and depth[0, 2] == -1
    assert ds.get_instance_length() == 3 and ds.get_occlusion_length() == 2 and ds.get_geometric_length() == 2
    assert ds.get_imgId_and_depth(1) == (0, "1=2")


def _cpu_rasterize(monkeypatch):
    monkeypatch.setattr(masks, "rasterize", lambda segms, hh, ww, device="cpu", stream=None:
                        torch.from_numpy(np.stack([M.decode_segm(s, hh, ww) for s in segms])))


def test_cocoa_reader(tmp_path, monkeypatch):
    """``COCOADataset`` (reference datasets/reader.py:209-291): regions -> modal / amodal masks, boxes from the masks,
    constant category 1, occlusion GT from the 1-based ``depth_constraint`` string with the 95 % occlusion rule."""
    _cpu_rasterize(monkeypatch)
    h, w = 50, 70
    polys = [[4, 4, 40, 6, 30, 40], [20, 10, 65, 12, 60, 45, 25, 35], [45, 25, 68, 25, 68, 48, 45, 48]]
    vis = M.decode_segm([polys[1]], h, w).copy()
    vis[:, :30] = 0
    regions = [dict(segmentation=polys[0], occlude_rate=0.0, isStuff=0),
               dict(segmentation=polys[1], occlude_rate=0.3, isStuff=0,
                    visible_mask=dict(size=[h, w], counts=M.rle_to_string(M.rle_encode(vis)).decode("ascii"))),
               dict(segmentation=polys[2], occlude_rate=0.97, isStuff=1)]
    data = dict(images=[dict(id=11, file_name="a.jpg", width=w, height=h)],
                annotations=[dict(regions=regions, depth_constraint="1-2,2-3,1-3")])
    p = tmp_path / "COCO_amodal_val2014.json"
    p.write_text(json.dumps(data))
    ds = reader.COCOADataset(str(p), device="cpu")
    assert ds.get_image_length() == 1 and ds.get_instance_length() == 3
    modal, category, bboxes, amodal, image_fn = ds.get_image_instances(0, with_gt=True)
    assert modal.shape == (3, h, w) and amodal.shape == (3, h, w) and image_fn == "a.jpg"
    assert np.array_equal(modal[1], vis) and np.array_equal(amodal[1], M.decode_segm([polys[1]], h, w))
    assert category.tolist() == [1, 1, 1]
    assert bboxes[1].tolist() == masks.mask_to_bbox(vis)
    gt = ds.get_gt_ordering(0)
    assert gt[0, 1] == 1 and gt[1, 2] == 0 and gt[0, 2] == 0 and gt.sum() == 1      # region 3 is > 95 % occluded
    assert ds.get_image_instances(0, ignore_stuff=True)[0].shape[0] == 2
    assert ds.get_image_instances(0, with_id=True)[-1] == 11
    m, b, c, fn, am = ds.get_instance(1, with_gt=True)
    assert np.array_equal(m, vis) and np.array_equal(am, amodal[1]) and c == 1


def test_kins_reader(tmp_path, monkeypatch):
    """``KINSLVISDataset`` (reference datasets/reader.py:460-539): annotations grouped by image in first-appearance
    order, inmodal RLE masks / boxes, amodal polygon masks."""
    _cpu_rasterize(monkeypatch)
    h, w = 40, 120
    polys = [[5, 5, 50, 5, 50, 30, 5, 30], [40, 10, 100, 10, 100, 35, 40, 35], [60, 2, 110, 2, 110, 20, 60, 20]]

    def ann(i, img, poly, cat):
        amodal = M.decode_segm([poly], h, w)
        inmodal = amodal.copy()
        inmodal[:, :w // 3] = 0
        return dict(id=i, image_id=img, category_id=cat, segmentation=[poly], inmodal_bbox=masks.mask_to_bbox(inmodal),
                    inmodal_seg=dict(size=[h, w], counts=M.rle_to_string(M.rle_encode(inmodal)).decode("ascii")))
    data = dict(images=[dict(id=2, file_name="b.png", width=w, height=h), dict(id=1, file_name="a.png", width=w, height=h)],
                annotations=[ann(0, 1, polys[0], 4), ann(1, 2, polys[1], 2), ann(2, 1, polys[2], 7)],
                categories=[dict(id=k) for k in range(8)])
    p = tmp_path / "instances_val.json"
    p.write_text(json.dumps(data))
    ds = reader.KINSLVISDataset("KINS", str(p), device="cpu")
    assert ds.get_image_length() == 2 and ds.get_instance_length() == 3 and ds.img_ids == [1, 2]
    modal, category, bboxes, amodal, image_fn = ds.get_image_instances(0, with_gt=True)
    assert image_fn == "a.png" and modal.shape == (2, h, w) and amodal.shape == (2, h, w)
    assert category.tolist() == [4, 7]
    assert np.array_equal(amodal[1], M.decode_segm([polys[2]], h, w))
    assert np.array_equal(modal[0], np.where(np.arange(w)[None] < w // 3, 0, amodal[0]))
    assert bboxes[0].tolist() == masks.mask_to_bbox(modal[0])
    assert ds.get_image_instances(1)[3].size == 0
