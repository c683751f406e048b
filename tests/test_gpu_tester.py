"""The evaluation driver (instaorder_b200/tester.py, mirror of the reference's tools/test.py::Tester) against the
reference's own per-image loop written with the reference-shaped API: same dataset-level numbers under the same keys."""
import types

import numpy as np
import pytest

from instaorder_b200 import inference, models, synth, tester
from oracle import calib, gen_golden

pytestmark = pytest.mark.gpu


class FakeReader(object):
    """The slice of the reference reader's interface the Tester uses (datasets/reader.py:335-457)."""

    def __init__(self, n_images, seed=0):
        rng = np.random.RandomState(seed)
        self.items = []
        for k in range(n_images):
            n = int(rng.randint(2, 7))
            image, masks, boxes = synth.make_scene(rng, 200, 260, n, wh_range=((30, 120), (30, 100)))
            occ, depth, overlap, count = synth.make_gt(rng, n)
            self.items.append(dict(image=image, masks=masks, boxes=boxes, occ=occ, depth=depth, overlap=overlap, count=count))

    def __len__(self):
        return len(self.items)

    def get_image_instances(self, i, with_gt=False):
        it = self.items[i]
        return it["masks"], np.ones(len(it["masks"]), np.int64), it["boxes"], np.array([]), "img_%d.jpg" % i

    def get_gt_ordering(self, i, kind, rm_bidirec=0):
        it = self.items[i]
        return [it["depth"], it["overlap"], it["count"]] if kind == "depth" else it["occ"]


def test_tester_matches_per_image_loop(golden_dir):
    case = "c2_od"
    c = gen_golden.CASES[case]
    params = dict(algo=c["algo"], backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=[2, 3]),
                  optim="SGD", lr=1e-4, weight_decay=1e-4, use_rgb=True, max_pairs=32)
    m = models.InstaOrderNet_od(params)
    m.load_state_dict(calib.load_calibrated(gen_golden.calib_path(case), c["wseed"], 5, [2, 3]))
    m.switch_to("eval")
    reader = FakeReader(7)
    args = types.SimpleNamespace(order_method="InstaOrderNet_od", pairs="all", zd=1, disp_select_method="",
                                 data=dict(patch_or_image="patch", input_size=256, remove_occ_bidirec=0,
                                           use_category=False), images_per_call=3)
    t = tester.Tester(args, m, reader, lambda fn: reader.items[int(fn.split("_")[1].split(".")[0])]["image"])
    out = t.run()
    # the reference's loop (tools/test.py:187-283) with the reference-shaped per-image API
    rec, pre, f1s, whdr = [], [], [], {k: [] for k in inference.WHDR_KEYS}
    for i in range(len(reader)):
        it = reader.items[i]
        occ, depth = inference.infer_order_sup_occ_depth(m, it["image"], it["masks"], t.expand_bbox(it["boxes"]), "all",
                                                         "InstaOrderNet_od", "patch", 256, "")
        w = inference.eval_depth_order_whdr(depth, [it["depth"], it["overlap"], it["count"]])
        for k in whdr:
            whdr[k].append(w[k][0])
        r, p, f = inference.eval_order_recall_precision_f1(occ, it["occ"], 1)
        rec.append(r); pre.append(p); f1s.append(f)
    assert out["val/recall"] == sum(rec) / len(rec) and out["val/precision"] == sum(pre) / len(pre)
    assert out["val/f1"] == sum(f1s) / len(f1s) and out["val/num_test_images"] == 7
    for k, vals in whdr.items():
        a = np.array(vals)
        mean = a[a != -1].sum() / (len(a[a != -1]) + 1e-6)
        ovl, eq = k.split("_")
        assert out["val_%s/WHDR_%s" % (ovl, eq)] == mean, k
