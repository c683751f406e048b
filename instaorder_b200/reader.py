"""``InstaOrderDataset`` with the interface of the reference's annotation reader (``datasets/reader.py:293-457``), on
top of this package's mask and GT-matrix producers: the InstaOrder annotation file (``annotations`` = per image
``image_id``, ``instance_ids``, ``occlusion``, ``depth``) + the COCO ``instances_{train,val}2017.json`` next to it
(located by the reference's own rule, :296-304).  No pycocotools / cvbase: plain ``json`` and a dict index; the modal
masks are rasterised on the GPU (``instaorder_b200.masks``), returned as numpy like the reference does or left in HBM
(``device_masks=True``: ``engine.Scene`` takes the CUDA tensor).  What the eval driver (``tester.Tester``) and the
training-side pair samplers need.  ``COCOADataset`` (reader.py:209-291) and ``KINSLVISDataset`` (reader.py:460-539)
-- the readers BASELINE config 3's evaluation goes through -- follow below with the same methods and return tuples
(``json`` instead of ``cvbase.load``)."""
import json
import os

import numpy as np

from . import annotations as _ann, masks as _masks


class InstaOrderDataset(object):
    def __init__(self, annot_fn, coco_annot_fn=None, device=None, device_masks=False):
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


def _stack(masks_list, h, w):
    return np.array(masks_list) if len(masks_list) else np.array([])


class COCOADataset(object):
    """reference datasets/reader.py:209-291: COCOA amodal annotations (``images`` + per-image ``regions`` with
    ``segmentation`` = amodal polygon, optional ``visible_mask`` RLE, ``occlude_rate``, ``isStuff``; the occlusion GT is
    the ``depth_constraint`` string "a-b,c-d": a occludes b)."""

    def __init__(self, annot_fn, device=None, device_masks=False):
        with open(annot_fn) as f:
            data = json.load(f)
        self.images_info = data["images"]
        self.annot_info = data["annotations"]
        self.indexing = [(i, j) for i, ann in enumerate(self.annot_info) for j in range(len(ann["regions"]))]
        self.device, self.device_masks = device, bool(device_masks)

    def __len__(self):
        return len(self.images_info)

    def get_instance_length(self):
        return len(self.indexing)

    def get_image_length(self):
        return len(self.images_info)

    def get_gt_ordering(self, imgidx):
        """reader.py:225-241: [idx1, idx2] = 1 for every "idx1-idx2" (1-based) unless idx2 is > 95 % occluded."""
        regions = self.annot_info[imgidx]["regions"]
        gt = np.zeros((len(regions), len(regions)), dtype=np.int64)
        order_str = self.annot_info[imgidx]["depth_constraint"]
        if len(order_str) == 0:
            return gt
        for o in order_str.split(","):
            a, b = o.split("-")
            a, b = int(a) - 1, int(b) - 1
            if regions[b]["occlude_rate"] > 0.95:
                continue
            gt[a, b] = 1
        return gt

    def _amodal(self, reg, h, w):
        return _masks.decode([reg["segmentation"]], h, w, self.device)

    def get_instance(self, idx, with_gt=False):
        imgidx, regidx = self.indexing[idx]
        img_info = self.images_info[imgidx]
        w, h = img_info["width"], img_info["height"]
        reg = self.annot_info[imgidx]["regions"][regidx]
        modal, bbox, category = _masks.read_COCOA(reg, h, w, self.device)
        return modal, bbox, category, img_info["file_name"], (self._amodal(reg, h, w) if with_gt else None)

    def get_image_instances(self, idx, with_id=False, with_gt=False, with_anns=False, ignore_stuff=False):
        ann_info, img_info = self.annot_info[idx], self.images_info[idx]
        image_fn, image_id = img_info["file_name"], img_info["id"]
        w, h = img_info["width"], img_info["height"]
        modal, bboxes, category, amodal = [], [], [], []
        for reg in ann_info["regions"]:
            if ignore_stuff and reg["isStuff"]:
                continue
            m, b, c = _masks.read_COCOA(reg, h, w, self.device)
            modal.append(m)
            bboxes.append(b)
            category.append(c)
            if with_gt:
                amodal.append(self._amodal(reg, h, w))
        ret = (_stack(modal, h, w), np.array(category), np.array(bboxes), _stack(amodal, h, w), image_fn)
        if with_anns:
            return ret + (ann_info, image_id)
        if with_id:
            return ret + (image_id,)
        return ret


class KINSLVISDataset(object):
    """reference datasets/reader.py:460-539: KINS (``inmodal_seg`` RLE + ``inmodal_bbox``, amodal ``segmentation``
    polygons) or LVIS-style annotations grouped by ``image_id`` in first-appearance order."""

    def __init__(self, dataset, annot_fn, device=None, device_masks=False):
        self.dataset = dataset
        with open(annot_fn) as f:
            data = json.load(f)
        self.images_info = data["images"]
        self.annot_info = data["annotations"]
        self.category_info = data.get("categories")
        self.imgfn_dict = dict((a["id"], a["file_name"]) for a in self.images_info)
        self.size_dict = dict((a["id"], (a["width"], a["height"])) for a in self.images_info)
        self.anns_dict = self.make_dict()
        self.img_ids = list(self.anns_dict.keys())
        self.device, self.device_masks = device, bool(device_masks)

    def __len__(self):
        return len(self.img_ids)

    def get_instance_length(self):
        return len(self.annot_info)

    def get_image_length(self):
        return len(self.img_ids)

    def make_dict(self):
        anns_dict = {}
        for ann in self.annot_info:
            anns_dict.setdefault(ann["image_id"], []).append(ann)
        return anns_dict

    def _read(self, ann, h, w):
        if self.dataset == "KINS":
            return _masks.read_KINS(ann, self.device)[:3]
        if self.dataset == "LVIS":
            return _masks.read_LVIS(ann, h, w, self.device)
        raise Exception("No such dataset: {}".format(self.dataset))

    def _amodal(self, ann, h, w):
        # decode(frPyObjects(polygons)) gives one [h, w] plane per polygon; .squeeze() of the single-polygon KINS
        # annotations is that plane (reader.py:489-491)
        return np.squeeze(np.stack([_masks.decode([p], h, w, self.device) for p in ann["segmentation"]], axis=-1))

    def get_instance(self, idx, with_gt=False):
        ann = self.annot_info[idx]
        w, h = self.size_dict[ann["image_id"]]
        modal, bbox, category = self._read(ann, h, w)
        return modal, bbox, category, self.imgfn_dict[ann["image_id"]], (self._amodal(ann, h, w) if with_gt else None)

    def get_image_instances(self, idx, with_gt=False, with_anns=False):
        imgid = self.img_ids[idx]
        w, h = self.size_dict[imgid]
        anns = self.anns_dict[imgid]
        modal, bboxes, category, amodal = [], [], [], []
        for ann in anns:
            m, b, c = self._read(ann, h, w)
            modal.append(m)
            bboxes.append(b)
            category.append(c)
            if with_gt:
                amodal.append(self._amodal(ann, h, w))
        ret = (_stack(modal, h, w), np.array(category), np.array(bboxes), _stack(amodal, h, w), self.imgfn_dict[imgid])
        return ret + (anns,) if with_anns else ret
