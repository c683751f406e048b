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
