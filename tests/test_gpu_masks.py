"""GPU rasteriser of COCO segmentations (io_masks_from_rle) against the CPU restatement of pycocotools' decode / merge
(bit-exact), through the reader-shaped wrappers, and as device-resident input of the order engine."""
import numpy as np
import pytest
import torch

from instaorder_b200 import engine, masks, synth
from oracle import coco_mask_oracle as M
from test_masks import random_polygon

pytestmark = pytest.mark.gpu


def make_segms(rng, h, w, n):
    segms = []
    for i in range(n):
        kind = i % 3
        if kind == 0:      # polygon list, 1-3 parts (merge = union)
            segms.append([random_polygon(rng, h, w, int(rng.randint(3, 20))) for _ in range(rng.randint(1, 4))])
        else:
            m = M.decode_segm([random_polygon(rng, h, w, int(rng.randint(3, 12)))], h, w)
            c = M.rle_encode(m)
            segms.append(dict(size=[h, w], counts=c if kind == 1 else M.rle_to_string(c)))
    return segms


@pytest.mark.parametrize("hw", [(480, 640), (375, 1242), (33, 17), (1, 9), (427, 640)])
def test_rasterize_matches_oracle(hw):
    h, w = hw
    rng = np.random.RandomState(h + w)
    segms = make_segms(rng, h, w, 11)
    got = masks.rasterize(segms, h, w).cpu().numpy()
    assert got.shape == (11, h, w) and got.dtype == np.uint8
    for i, s in enumerate(segms):
        assert np.array_equal(got[i], M.decode_segm(s, h, w)), "instance %d" % i
    assert masks.rasterize([], h, w).shape == (0, h, w)


def test_reader_wrappers():
    rng = np.random.RandomState(3)
    h, w = 120, 160
    poly = random_polygon(rng, h, w, 9)
    want = M.decode_segm([poly], h, w)
    m, bbox, cat = masks.read_LVIS(dict(segmentation=[poly], bbox=[1, 2, 3, 4], category_id=7), h, w)
    assert np.array_equal(m, want) and bbox == [1, 2, 3, 4] and cat == 7
    rle = dict(size=[h, w], counts=M.rle_to_string(M.rle_encode(want)))
    m, bbox, cat, score = masks.read_KINS(dict(inmodal_seg=rle, inmodal_bbox=[5, 6, 7, 8], category_id=2))
    assert np.array_equal(m, want) and score == 1.
    m, bbox, cat = masks.read_COCOA(dict(segmentation=poly), h, w)
    assert np.array_equal(m, want) and bbox == masks.mask_to_bbox(want) and cat == 1
    with pytest.raises(ValueError):
        masks.rasterize([dict(size=[h, w], counts=[5, 5])], h, w)


def test_device_masks_feed_the_engine():
    """Masks rasterised on the GPU go into the order engine without a host round trip; results equal the host path."""
    rng = np.random.RandomState(11)
    h, w, n = 200, 260, 5
    image = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    segms = [[random_polygon(rng, h, w, 8)] for _ in range(n)]
    d_masks = masks.rasterize(segms, h, w)
    h_masks = d_masks.cpu().numpy()
    boxes = np.asarray([masks.mask_to_bbox(m) for m in h_masks], dtype=np.float64)
    assert (boxes[:, 2] > 0).all()
    bexp = engine.expand_bbox(boxes, 3.0)
    eng = engine.OrderEngine([2, 3], 256, max_pairs=16)
    eng.load_state_dict(synth.random_state_dict(1, 5, [2, 3]))
    r_host = eng.infer_scenes([engine.Scene(image, h_masks, bexp)], "InstaOrderNet_od", "all", "patch", return_details=True)[0]
    h2d_host = eng.h2d_bytes
    r_dev = eng.infer_scenes([engine.Scene(image, d_masks, bexp)], "InstaOrderNet_od", "all", "patch", return_details=True)[0]
    h2d_dev = eng.h2d_bytes - h2d_host
    assert np.array_equal(r_host["logits"], r_dev["logits"])
    assert np.array_equal(r_host["occ"], r_dev["occ"]) and np.array_equal(r_host["depth"], r_dev["depth"])
    assert h2d_dev <= h2d_host - n * h * w
    torch.cuda.synchronize()
