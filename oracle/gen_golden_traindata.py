"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the UNMODIFIED reference training datasets' ``__getitem__``
(datasets/depth_occ_order_dataset.py, depth_order_dataset.py, occ_order_dataset.py; ``patch`` mode) on a synthetic
scene served by a mocked annotation reader into tests/golden/traindata.npz.
Run in the build container:  ``python -m oracle.gen_golden_traindata``."""
import os
import sys

import numpy as np

from instaorder_b200 import synth
from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SZ = 64
BASE_AUG = dict(flip=True, shift=[-0.2, 0.2], scale=[0.8, 1.2])
N_SAMPLES = 12
SCENE = dict(seed=77, H=200, W=260, N=5, wh_range=((30, 120), (30, 100)))


def make_scene():
    rng = np.random.RandomState(SCENE["seed"])
    image, masks, boxes = synth.make_scene(rng, SCENE["H"], SCENE["W"], SCENE["N"], wh_range=SCENE["wh_range"],
                                           float_boxes=True)
    occ, depth, overlap, count = synth.make_gt(rng, SCENE["N"])
    geo = []      # the "i<j" / "i=j" strings of reader.get_imgId_and_depth
    for i in range(SCENE["N"]):
        for j in range(SCENE["N"]):
            if i != j and depth[i, j] == 1 and depth[j, i] == 0:
                geo.append("%d<%d" % (i, j))
            elif i < j and depth[i, j] == 2:
                geo.append("%d=%d" % (i, j))
    return image, masks, boxes, occ, depth, overlap, count, geo


class MockReader(object):
    def __init__(self, scene):
        self.image, self.masks, self.boxes, self.occ, self.depth, self.overlap, self.count, self.geo = scene

    def get_image_instances(self, idx, with_gt=True):
        return self.masks.copy(), np.ones(len(self.masks), np.int64), self.boxes.copy(), self.masks.copy(), "img.jpg"

    def get_gt_ordering(self, idx, type="occlusion", rm_bidirec=0, rm_overlap=0):
        if type == "depth":
            return self.depth.copy(), self.overlap.copy(), self.count.copy()
        return self.occ.copy()

    def get_imgId_and_depth(self, idx):
        return 0, self.geo[idx % len(self.geo)]

    def get_geometric_length(self):
        return len(self.geo)

    def get_instance_length(self):
        return 1


def load_reference_datasets():
    ns = ref_shim.load()
    for k in list(sys.modules):
        if k.startswith("_instaorder_ref."):
            sys.modules.setdefault(k[len("_instaorder_ref."):], sys.modules[k])
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    try:
        import datasets as r_datasets
    finally:
        sys.path.remove(ref_shim.REFERENCE_ROOT)
    return r_datasets


def make_dataset(r_datasets, cls_name, algo, scene, mode="patch"):
    import torchvision.transforms as transforms
    from PIL import Image
    cls = getattr(r_datasets, cls_name)
    ds = object.__new__(cls)
    ds.algo = algo
    ds.dataset = "InstaOrder"
    ds.rm_bidirec = 0
    ds.rm_overlap = 0
    ds.data_reader = MockReader(scene)
    ds.img_transform = transforms.Compose([transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    ds.sz = SZ
    ds.phase = "train"
    ds.config = dict(base_aug=BASE_AUG, load_rgb=True, patch_or_image=mode, train_image_root="", use_category=False,
                     extend_bidirec=False, dataset="InstaOrder")
    ds.memcached = False
    ds.get_pair_patch_or_image = {"patch": ds._get_pair, "resize": ds._get_pair_resize,
                                  "image": ds._get_pair_image}[mode]
    ds._load_image = lambda fn: Image.fromarray(scene[0])
    return ds


def main():
    r_datasets = load_reference_datasets()
    scene = make_scene()
    out = {}
    for name, cls_name, algo in (("od", "SupDepthOccOrderDataset", "InstaOrderNet_od"),
                                 ("d", "SupDepthOrderDataset", "InstaOrderNet_d"),
                                 ("o", "SupOcclusionOrderDataset", "InstaOrderNet_o"),
                                 ("ordernet", "SupOcclusionOrderDataset", "OrderNet")):
        ds = make_dataset(r_datasets, cls_name, algo, scene)
        for k in range(N_SAMPLES):
            np.random.seed(1000 + k)
            s = ds[k]
            x = np.concatenate([s[1].numpy(), s[2].numpy(), s[0].numpy()], 0).astype(np.float32)   # (m1, m2, rgb)
            out["%s_%d_x" % (name, k)] = x
            out["%s_%d_labels" % (name, k)] = np.concatenate(
                [np.asarray(v, dtype=np.float64).reshape(-1) for v in s[3:]])
    # the shipped ^od / ^d training configs use patch_or_image = "resize"; "image" is the third branch
    for mode in ("resize", "image"):
        ds = make_dataset(r_datasets, "SupDepthOccOrderDataset", "InstaOrderNet_od", scene, mode)
        for k in range(N_SAMPLES):
            np.random.seed(2000 + k)
            s = ds[k]
            out["od_%s_%d_x" % (mode, k)] = np.concatenate([s[1].numpy(), s[2].numpy(), s[0].numpy()], 0).astype(np.float32)
            out["od_%s_%d_labels" % (mode, k)] = np.concatenate(
                [np.asarray(v, dtype=np.float64).reshape(-1) for v in s[3:]])
    path = os.path.join(GOLDEN, "traindata.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
