"""G13: the training-side ``__getitem__`` of the reference datasets, ``patch`` mode.
CPU: the numpy restatement (oracle/train_data_oracle.py) and the product's host-side sampler
(instaorder_b200/train_data.py: same np.random draw order, crop geometry, labels) against the fixtures produced by the
UNMODIFIED reference dataset classes (oracle/gen_golden_traindata.py) -- bit-exact tensors and labels."""
import os

import numpy as np
import pytest

from instaorder_b200 import train_data as TD
from oracle import gen_golden_traindata as GG
from oracle import train_data_oracle as TO

CASES = [("od", "InstaOrderNet_od"), ("d", "InstaOrderNet_d"), ("o", "InstaOrderNet_o"), ("ordernet", "OrderNet")]


def _oracle_sample(name, algo, scene, k):
    image, masks, boxes, occ, depth, overlap, count, geo = scene
    np.random.seed(1000 + k)
    if name in ("od", "d"):
        s = geo[k % len(geo)]
        i1, i2 = map(int, s.split("<" if "<" in s else "="))
        if name == "od":
            r = TO.getitem_od(image, masks, boxes, i1, i2, depth, overlap, count, occ, GG.SZ, GG.BASE_AUG)
            labels = [r["depth"], r["count"], r["overlap"]] + list(r["occ"])
        else:
            r = TO.getitem_d(image, masks, boxes, i1, i2, depth, overlap, count, GG.SZ, GG.BASE_AUG)
            labels = [r["depth"], r["count"], r["overlap"]]
    else:
        r = TO.getitem_occ(algo, image, masks, boxes, occ, GG.SZ, GG.BASE_AUG)
        labels = [r["label"]] if algo == "OrderNet" else list(r["occ"])
    x = np.concatenate([r["modal1"][None].astype(np.float32), r["modal2"][None].astype(np.float32), r["rgb"]], 0)
    return x, np.asarray(labels, dtype=np.float64), r


@pytest.mark.parametrize("name,algo", CASES)
def test_oracle_getitem_matches_reference(name, algo, golden_dir):
    z = np.load(os.path.join(golden_dir, "traindata.npz"))
    scene = GG.make_scene()
    for k in range(GG.N_SAMPLES):
        x, labels, _ = _oracle_sample(name, algo, scene, k)
        assert np.array_equal(labels, z["%s_%d_labels" % (name, k)]), (name, k, labels, z["%s_%d_labels" % (name, k)])
        assert np.array_equal(x[:2], z["%s_%d_x" % (name, k)][:2]), "masks of %s sample %d" % (name, k)
        assert np.array_equal(x[2:], z["%s_%d_x" % (name, k)][2:]), "rgb of %s sample %d" % (name, k)


@pytest.mark.parametrize("name,algo", CASES)
def test_host_sampler_matches_oracle(name, algo):
    """instaorder_b200.train_data.sample_pair: same draws, crop box, flip, swap and labels as the oracle."""
    scene = GG.make_scene()
    image, masks, boxes, occ, depth, overlap, count, geo = scene
    gt = dict(occ=occ, depth=depth, overlap=overlap, count=count)
    for k in range(GG.N_SAMPLES):
        _, labels, r = _oracle_sample(name, algo, scene, k)
        np.random.seed(1000 + k)
        pair = None
        if name in ("od", "d"):
            s = geo[k % len(geo)]
            pair = tuple(map(int, s.split("<" if "<" in s else "=")))
        spec = TD.sample_pair(algo, boxes, gt, GG.BASE_AUG, pair=pair)
        assert [spec.x, spec.y, spec.s, spec.s] == list(r["new_bbox"]) and spec.flip == r["flip"]
        assert spec.swapped == r["swapped"]
        assert np.array_equal(np.asarray(spec.labels, dtype=np.float64), labels)
        if pair is None:
            assert (spec.idx1, spec.idx2) == r["idx"]


@pytest.mark.parametrize("mode", ["resize", "image"])
def test_oracle_getitem_whole_image_modes(mode, golden_dir):
    """`resize` (the shipped ^od / ^d training config) and `image` modes: u8 INTER_LINEAR + nearest + flip + swap."""
    z = np.load(os.path.join(golden_dir, "traindata.npz"))
    image, masks, boxes, occ, depth, overlap, count, geo = GG.make_scene()
    gt = dict(occ=occ, depth=depth, overlap=overlap, count=count)
    for k in range(GG.N_SAMPLES):
        s = geo[k % len(geo)]
        i1, i2 = map(int, s.split("<" if "<" in s else "="))
        np.random.seed(2000 + k)
        r = TO.getitem_od(image, masks, boxes, i1, i2, depth, overlap, count, occ, GG.SZ, GG.BASE_AUG, mode=mode)
        x = np.concatenate([r["modal1"][None].astype(np.float32), r["modal2"][None].astype(np.float32), r["rgb"]], 0)
        assert np.array_equal(x, z["od_%s_%d_x" % (mode, k)]), (mode, k)
        labels = np.asarray([r["depth"], r["count"], r["overlap"]] + list(r["occ"]), dtype=np.float64)
        assert np.array_equal(labels, z["od_%s_%d_labels" % (mode, k)])
        np.random.seed(2000 + k)
        spec = TD.sample_pair("InstaOrderNet_od", boxes, gt, GG.BASE_AUG, pair=(i1, i2), mode=mode)
        assert spec.flip == r["flip"] and spec.swapped == r["swapped"]
        assert np.array_equal(np.asarray(spec.labels, dtype=np.float64), labels)
