"""``InstaOrderDataset`` with the interface of the reference's annotation reader (``datasets/reader.py:293-457``), on
top of this package's mask and GT-matrix producers: the InstaOrder annotation file (``annotations`` = per image
``image_id``, ``instance_ids``, ``occlusion``, ``depth``) + the COCO ``instances_{train,val}2017.json`` next to it
(located by the reference's own rule, :296-304).  No pycocotools / cvbase: plain ``json`` and a dict index; the modal
masks are rasterised on the GPU (``instaorder_b200.masks``), returned as numpy like the reference does or left in HBM
(``device_masks=True``: ``engine.Scene`` takes the CUDA tensor).  What the eval driver (``tester.Tester``) and the
training-side pair samplers need, nothing else of that file (KINS / COCOA dataset classes are not mirrored)."""
import json
import os

import numpy as np

from . import annotations as _ann, masks as _masks


class InstaOrderDataset(object):
    def __init__(self, annot_fn, coco_annot_fn=None, device="cuda:0", device_masks=False):
        with open(annot_fn) as f:
            data = json.load(f)
        self.annot_info = data["annotations"]
        if coco_annot_fn is None:                     # reader.py:296-304
            data_type = None
            for dtype in ("train2017", "val2017"):
                if dtype in annot_fn:
                    data_type = dtype
            coco_annot_fn = os.path.join(os.path.dirname(annot_fn), "instances_%s.json" % data_type)
        with open(coco_annot_fn) as f:
            coco = json.load(f)
        self._imgs = {im["id"]: im for im in coco["images"]}
        self._anns = {a["id"]: a for a in coco["annotations"]}
        self.device, self.device_masks = device, bool(device_masks)

    def __len__(self):
        return len(self.annot_info)

    # ---- index helpers used by the training datasets (reader.py:306-333) ------------------------------------------
    def get_image_length(self):
        return len(self.annot_info)

    def get_instance_length(self):
        self.indexing = [(i, k) for i, a in enumerate(self.annot_info) for k in range(len(a["instance_ids"]))]
        return len(self.indexing)

    def get_occlusion_length(self):
        self.occ_all_img_and_idx = [(i, k) for i, a in enumerate(self.annot_info) for k in range(len(a["occlusion"]))]
        return len(self.occ_all_img_and_idx)

    def get_geometric_length(self):
        self.depth_all_img_and_order = [(i, g["order"]) for i, a in enumerate(self.annot_info) for g in a["depth"]]
        return len(self.depth_all_img_and_order)

    def get_imgId_and_depth(self, depth_all_idx):
        return self.depth_all_img_and_order[depth_all_idx]

    # ---- ground truth (reader.py:335-400) ----------------------------------------------------------------------------
    def get_gt_ordering(self, imgidx, type, rm_bidirec=0, rm_overlap=0):
        return _ann.gt_ordering(self.annot_info[imgidx], type, rm_bidirec, rm_overlap)

    # ---- instances of one image (reader.py:421-457) ----------------------------------------------------------------
    def get_image_instances(self, idx, with_id=False, with_gt=False, with_anns=False, ignore_stuff=False):
        ann_info = self.annot_info[idx]
        image_id = ann_info["image_id"]
        img_info = self._imgs[image_id]
        image_fn = img_info["file_name"]
        w, h = img_info["width"], img_info["height"]
        anns = [self._anns[int(a)] for a in ann_info["instance_ids"]]
        modal, bboxes, category = _masks.image_instances(anns, h, w, self.device)
        if not self.device_masks:
            modal = modal.cpu().numpy()
        amodal = np.array([])                               # the reference returns an empty array here too (:440-443)
        if with_anns:
            return modal, category, bboxes, amodal, image_fn, ann_info, image_id
        if with_id:
            return modal, category, bboxes, amodal, image_fn, image_id
        return modal, category, bboxes, amodal, image_fn

    def get_instance(self, idx, with_gt=False):
        """reader.py:402-419 (one region; call get_instance_length() first, as the reference requires)."""
        imgidx, regidx = self.indexing[idx]
        img_info = self._imgs[self.annot_info[imgidx]["image_id"]]
        ann = self._anns[int(self.annot_info[imgidx]["instance_ids"][regidx])]
        modal, bbox, category = _masks.read_LVIS(ann, img_info["height"], img_info["width"], self.device)
        return modal, bbox, category, img_info["file_name"], None
