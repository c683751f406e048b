"""Fused pair gather (CUDA) vs the oracle / reference fixtures.  Integer + byte work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from instaorder_b200 import _lib, engine, synth
from oracle import gen_golden, oracle as O
import gpu_util as U

pytestmark = pytest.mark.gpu


def run_gather(scene, boxes, D, pairs=None, mode="patch"):
    image, masks, _ = scene
    sc = engine.Scene(image, masks, boxes)
    pr = engine.enumerate_pairs(sc.n) if pairs is None else np.asarray(pairs, dtype=np.int32)
    eng = run_gather.engines.get(D)
    if eng is None:
        eng = run_gather.engines[D] = engine.OrderEngine([2, 3], D, max_pairs=64)
    crops = engine.pair_crop_boxes(sc.boxes, pr) if mode == "patch" else None
    s, P = eng.stage_batch([(sc, pr, crops, 0, 0)], mode)
    eng.gather(s, P, mode)
    torch.cuda.synchronize()
    inner, raw = U.unpack_pair_tensor(eng.pair_tensor, P, D)
    return pr, crops, inner, raw


run_gather.engines = {}


def expected_patch(image, masks, boxes, i, j, D):
    rgb, mi, mj, nb = O.pair_patch(image, masks, boxes, i, j, D)
    return U.f32_to_bf16_rn(O.pair_tensor(rgb, mi, mj)), nb


@pytest.mark.parametrize("case", ["c1_o", "c2_od", "c3_ordernet"])
def test_gather_patch_matches_reference(golden_dir, case):
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    scene = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(scene[2], 3.0)
    assert np.array_equal(bexp, z["boxes_expanded"])
    pr, crops, inner, raw = run_gather(scene, bexp, 256)
    P = pr.shape[0]
    # borders and the 3 padding channels are zero
    assert not raw[:, :3].any() and not raw[:, -3:].any() and not raw[:, :, :3].any() and not raw[:, :, 259:].any()
    assert not raw[..., 5:].any()
    for k in range(P):
        # masks: bit-exact against the digests of what the *reference* fed its network
        assert gen_golden.digest(inner[k, :2].astype(np.uint8)) == z["mask_digest"][k], (case, k)
    for n, k in enumerate(z["full_idx"]):
        want = U.f32_to_bf16_rn(z["full_x"][n])          # bf16(reference fp32 tensor)
        assert np.array_equal(inner[k], want), (case, int(k), float(np.abs(inner[k] - want).max()))
    # every pair against the oracle (which test_oracle_golden pins to the reference digests)
    for k, (i, j) in enumerate(pr):
        want, nb = expected_patch(scene[0], scene[1], bexp, int(i), int(j), 256)
        assert list(crops[k]) == nb
        assert np.array_equal(inner[k], want), (case, k)


def test_gather_patch_edge_cases():
    """Crops far outside the image, 1-pixel crops, heavy up/down-scaling, non-256 output sizes."""
    rng = np.random.RandomState(11)
    image = rng.randint(0, 256, size=(97, 131, 3)).astype(np.uint8)
    masks = (rng.rand(4, 97, 131) < 0.5).astype(np.uint8)
    masks[3] *= 7                                  # use_category-style mask values (reference occ_order_dataset.py:183)
    boxes = np.array([[-400.0, -300.0, 50.0, 40.0],     # entirely outside -> all padding
                      [10.2, 20.7, 1.0, 1.0],           # tiny box -> S small (upscaling x100+)
                      [0.0, 0.0, 131.0, 97.0],          # whole image
                      [100.0, 80.0, 900.0, 700.0]])     # huge -> strong downscaling
    for D in (64, 256):
        pr, crops, inner, raw = run_gather((image, masks, None), boxes, D)
        for k, (i, j) in enumerate(pr):
            want, nb = expected_patch(image, masks, boxes, int(i), int(j), D)
            assert list(crops[k]) == nb
            assert np.array_equal(inner[k], want), (D, k, nb)


def test_gather_degenerate_pair_is_an_error():
    boxes = np.zeros((2, 4))          # two empty masks -> mask_to_bbox gives [0,0,0,0] -> int(size) == 0
    with pytest.raises(_lib.IoError) as e:
        engine.pair_crop_boxes(boxes, [[0, 1]])
    assert e.value.code == _lib.IO_ERR_DEGENERATE


def test_gather_resize_mode(golden_dir):
    case = "c2_od_resize"
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    scene = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(scene[2], 3.0)
    pr, _, inner, raw = run_gather(scene, bexp, 384, mode="resize")
    for k in range(pr.shape[0]):
        assert gen_golden.digest(inner[k, :2].astype(np.uint8)) == z["mask_digest"][k], k
    k = int(z["full_idx"][0])
    ref = z["full_x"][0]
    # float64 cubic without u8 rounding: tolerance = one bf16 ulp of the normalised value (|v| < 4 -> 2^-6)
    assert np.abs(inner[k, 2:] - ref[2:]).max() <= 2.0 ** -6
    assert np.abs(inner[k, 2:] - U.f32_to_bf16_rn(ref[2:])).mean() < 1e-4
    assert not raw[:, :3].any() and not raw[:, -3:].any() and not raw[:, :, :3].any() and not raw[:, :, 387:].any()


def test_bordering_kernel():
    rng = np.random.RandomState(5)
    image, masks, boxes = synth.make_scene(rng, 120, 160, 7, wh_range=((10, 80), (10, 60)))
    masks[2] *= 3                     # a category-valued mask: (dilate == 1) is then never true for it
    sc = engine.Scene(image, masks, boxes)
    eng = run_gather.engines.get(256) or engine.OrderEngine([2, 3], 256, max_pairs=64)
    run_gather.engines[256] = eng
    pr = engine.enumerate_pairs(7)
    got = eng.bordering(sc, pr)
    want = np.array([O.bordering(masks[i], masks[j]) for i, j in pr])
    assert np.array_equal(got, want)
    assert want.any() and not want.all()


def test_gather_image_mode(golden_dir):
    """`image` mode: padded-square INTER_LINEAR rgb (bit-exact) + nearest masks over the padded square."""
    case = "c1_o_image"
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    scene = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(scene[2], 3.0)
    pr, _, inner, raw = run_gather(scene, bexp, 256, mode="image")
    for k in range(pr.shape[0]):
        assert gen_golden.digest(inner[k, :2].astype(np.uint8)) == z["mask_digest"][k], k
    k = int(z["full_idx"][0])
    assert np.array_equal(inner[k], U.f32_to_bf16_rn(z["full_x"][0]))
    want = U.f32_to_bf16_rn(O.image_mode_rgb(scene[0], 256))
    for k in range(pr.shape[0]):
        assert np.array_equal(inner[k, 2:], want)
