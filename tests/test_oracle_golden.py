"""Pins oracle/oracle.py (the CPU restatement) against fixtures frozen from the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import calib, gen_golden, oracle as O


def test_geometry_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "geometry.npz"))
    for b, nb in zip(z["boxes"], z["crops"]):
        assert O.pair_crop_box(b, 0, 1) == list(nb)
        bi = b.astype(np.int64)
        if np.array_equal(bi, b):      # integer boxes must give the same geometry as their float64 image
            assert O.pair_crop_box(bi, 0, 1) == list(nb)
    for roi, rgb, m in zip(z["rois"], z["rgb"], z["m"]):
        got = O.resize_cubic_u8(O.crop_padding(z["image"], roi, 0), 64, 64)
        assert np.array_equal(got, rgb), roi
        assert np.array_equal(O.resize_nearest(O.crop_padding(z["mask"], roi, 0), 64, 64), m), roi


def test_metrics_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    for t in range(len(z["N"])):
        n = int(z["N"][t])
        prf = O.eval_order_recall_precision_f1(z["order"][t][:n, :n], z["gt"][t][:n, :n], int(z["zd"][t]))
        assert tuple(prf) == tuple(z["prf"][t]), t          # bit-identical float64
        wh = O.eval_depth_order_whdr(z["depth_pred"][t][:n, :n],
                                     (z["gtd"][t][:n, :n], z["ovl"][t][:n, :n], z["cnt"][t][:n, :n]))
        got = np.array([float(wh[k][0]) for k in O.WHDR_KEYS])
        assert np.array_equal(got, z["whdr"][t]), t


@pytest.mark.parametrize("case", ["c1_o", "c2_od", "c3_ordernet", "c2_d", "c3_ordernet_ext", "c2_od_resize",
                                  "c1_o_image", "c2_od_full", "c3_ordernet_full", "c2_od_big"])
def test_order_golden(golden_dir, case):
    c = gen_golden.CASES[case]
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    image, masks, boxes = gen_golden.build_scene(case)
    bexp = O.expand_bbox(boxes, 3.0)
    assert np.array_equal(bexp, z["boxes_expanded"])
    mode = c.get("patch_or_image", "patch")
    D = c.get("input_size", 256)
    N = masks.shape[0]
    plist = O.enumerate_pairs(N)
    # gather: masks bit-exact for every pair, rgb fp32 tensor bit-exact (patch) / 1e-5 (resize: float64 cubic)
    rgb_whole = O.resize_mode_rgb(image, D) if mode == "resize" else \
        (O.image_mode_rgb(image, D) if mode == "image" else None)
    for k, (i, j) in enumerate(plist):
        if mode == "patch":
            rgb, mi, mj, _ = O.pair_patch(image, masks, bexp, i, j, D)
            x = O.pair_tensor(rgb, mi, mj)
            assert gen_golden.digest(x[2:]) == z["rgb_digest"][k], (case, k)
        elif mode == "image":
            x = np.concatenate([O.image_mode_mask(masks[i], D)[None].astype(np.float32),
                                O.image_mode_mask(masks[j], D)[None].astype(np.float32), rgb_whole])
            assert gen_golden.digest(x[2:]) == z["rgb_digest"][k], (case, k)
        else:
            x = np.concatenate([O.resize_mode_mask(masks[i], D)[None].astype(np.float32),
                                O.resize_mode_mask(masks[j], D)[None].astype(np.float32), rgb_whole])
        assert gen_golden.digest(x[:2].astype(np.uint8)) == z["mask_digest"][k], (case, k)
        if k in list(z["full_idx"]):
            ref = z["full_x"][list(z["full_idx"]).index(k)]
            assert np.array_equal(x[:2], ref[:2])
            if mode in ("patch", "image"):
                assert np.array_equal(x, ref)
            else:
                assert np.abs(x - ref).max() < 2e-5
    # network + decisions
    sd = gen_golden.state_dict_for(case)
    r = O.infer_order(sd, image, masks, bexp, "all", c["algo"], mode, D)
    heads = ["fc_occ", "fc_depth"] if c["algo"] == "InstaOrderNet_od" else ["fc"]
    for h, head in enumerate(heads):
        got = np.stack([np.stack(r["logits"][p][head]) for p in plist])
        # batched vs batch-1 fp32 conv; the tolerance follows the head scale of the case
        assert np.abs(got - z["logits%d" % h]).max() < 2e-4 * c.get("head_scale", 1.0), (case, head)
    if c["algo"] != "InstaOrderNet_d":
        ok = r["margin_occ"] > 1e-3
        assert np.array_equal(r["occ"][ok], z["occ"][ok])
        assert ok.sum() >= 0.8 * N * (N - 1)
    if c["algo"] in ("InstaOrderNet_od", "InstaOrderNet_d"):
        ok = r["margin_depth"] > 1e-3
        assert np.array_equal(r["depth"][ok], z["depth"][ok])


def test_bordering_matches_definition():
    rng = np.random.RandomState(3)
    a = np.zeros((12, 12), np.uint8); b = np.zeros((12, 12), np.uint8)
    a[2:5, 2:5] = 1
    b[5:7, 2:5] = 1            # touches below
    assert O.bordering(a, b)
    b[:] = 0; b[6:8, 2:5] = 1  # one pixel gap
    assert not O.bordering(a, b)
    b[:] = 0; b[5, 5] = 1      # diagonal only: the cross kernel does not reach it
    assert not O.bordering(a, b)
    a[:] = 0; a[0, 0] = 1; b[:] = 0; b[0, 1] = 1
    assert O.bordering(a, b)


def test_gt_order_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "gt_order.npz"))
    for t in range(3):
        modal, amodal = gen_golden.kins_scene(int(z["seed%d" % t]))
        assert np.array_equal(O.infer_gt_order(modal, amodal), z["gt%d" % t])
        assert z["gt%d" % t].sum() > 0


def test_orig_mode_oracle_matches_reference(golden_dir):
    """``patch_or_image='orig'`` (reference inference.py:401-408): H, W rounded to multiples of 32, a
    non-square network input.  Oracle restatement vs the unmodified reference's logits / matrices (fixture from
    oracle/gen_golden_orig.py).  The CUDA path does not run this mode yet -- groundwork."""
    from oracle import gen_golden_orig
    z = np.load(os.path.join(golden_dir, "order_c2_od_orig.npz"))
    c = gen_golden.CASES[gen_golden_orig.CASE]
    image, masks, boxes = gen_golden.build_scene(gen_golden_orig.CASE)
    assert tuple(z["net_input_shape"]) == (1, 5, O.get_closest_int_multiple_of(image.shape[0], 32),
                                           O.get_closest_int_multiple_of(image.shape[1], 32))
    assert O.get_closest_int_multiple_of(48, 32) == 64 and O.get_closest_int_multiple_of(47, 32) == 32
    sd = calib.load_calibrated(gen_golden.calib_path(gen_golden_orig.CASE), c["wseed"], 5, c["num_classes"])
    r = O.infer_order(sd, image, masks, boxes, "all", "InstaOrderNet_od", "orig", c["input_size"])
    plist = r["pairs"]
    for h, head in enumerate(("fc_occ", "fc_depth")):
        got = np.stack([np.stack(r["logits"][p][head]) for p in plist])
        assert np.abs(got - z["logits%d" % h]).max() < 2e-4
    ok = r["margin_occ"] > 1e-3
    assert np.array_equal(r["occ"][ok], z["occ"][ok])
    ok = r["margin_depth"] > 1e-3
    assert np.array_equal(r["depth"][ok], z["depth"][ok])
