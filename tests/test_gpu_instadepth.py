"""InstaDepthNet^od order inference (BASELINE config 5) on the GPU against the fixture frozen from the UNMODIFIED
reference (``infer_order_sup_occ_depth(method='InstaDepthNet_od')``, resize 384^2): logits of both heads within the
north_star tolerance (2e-2 absolute, bf16 vs the fp32 reference), order matrices identical off ties, and against the
ideal-bf16 CPU emulation of the same arithmetic (kernel-correctness check proper)."""
import os

import numpy as np
import pytest

from instaorder_b200 import engine, inference, models
from instaorder_b200.depth_engine import DepthOrderEngine
from oracle import gen_golden_instadepth as G, instadepth_oracle as IO, oracle as O

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
MARGIN = 5e-3      # decisions are compared where the reference's probability margin exceeds the bf16 logit noise / 4


@pytest.fixture(scope="module")
def setup(golden_dir):
    z = np.load(os.path.join(golden_dir, "instadepth_order.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED)
    eng = DepthOrderEngine(G.D, max_pairs=16, max_images=4)
    eng.load_state_dict(sd)
    return z, sd, eng


def test_logits_and_matrices_match_reference(setup):
    z, sd, eng = setup
    image, masks, boxes = G.build_scene()
    r = eng.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od", "all", "resize", return_details=True)[0]
    occ_l, depth_l = r["logits"][:, :, 0:2], r["logits"][:, :, 2:5]
    e_o, e_d = np.abs(occ_l - z["occ_logits"]).max(), np.abs(depth_l - z["depth_logits"]).max()
    print("InstaDepthNet_od: max |logit - reference fp32| occ %.5f depth %.5f" % (e_o, e_d))
    assert e_o < LOGIT_TOL and e_d < LOGIT_TOL
    n = G.N_INST
    checked = 0
    for p, (i, j) in enumerate(O.enumerate_pairs(n)):
        _, m_d = O.decide_depth(z["depth_logits"][p, 0], z["depth_logits"][p, 1])
        _, _, m_o = O.decide_occ(z["occ_logits"][p, 0], z["occ_logits"][p, 1])
        if m_d > MARGIN:
            assert r["depth"][i, j] == z["depth"][i, j] and r["depth"][j, i] == z["depth"][j, i]
            checked += 1
        if m_o > MARGIN:
            assert r["occ"][i, j] == z["occ"][i, j] and r["occ"][j, i] == z["occ"][j, i]
            checked += 1
    assert checked >= 6


def test_logits_match_ideal_bf16_emulation(setup):
    z, sd, eng = setup
    image, masks, boxes = G.build_scene()
    r = eng.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od", "all", "resize", return_details=True)[0]
    rgb = O.resize_mode_rgb(image, G.D)[None]
    mm = [O.resize_mode_mask(m, G.D)[None].astype(np.float32) for m in masks]
    m1, m2 = [], []
    for (i, j) in O.enumerate_pairs(G.N_INST):
        m1 += [mm[i], mm[j]]
        m2 += [mm[j], mm[i]]
    P = len(m1) // 2
    emu = IO.order_forward(sd, rgb, np.stack(m1), np.stack(m2), np.zeros(2 * P, np.int64), bf16=True)
    e_d = np.abs(r["logits"][:, :, 2:5] - emu["depth"].reshape(P, 2, 3)).max()
    e_o = np.abs(r["logits"][:, :, 0:2] - emu["occ"].reshape(P, 2, 2)).max()
    print("InstaDepthNet_od: max |logit - ideal bf16 emulation| occ %.5f depth %.5f" % (e_o, e_d))
    assert e_o < 1e-2 and e_d < 1e-2


def test_batching_and_reference_api(setup, golden_dir):
    """Two images in one batch (encoder features broadcast by index) = each image alone; the reference-shaped API."""
    z, sd, eng = setup
    from instaorder_b200 import synth
    image, masks, boxes = G.build_scene()
    rng = np.random.RandomState(9)
    img2, masks2, boxes2 = synth.make_scene(rng, 333, 500, 3, wh_range=((50, 200), (50, 200)))
    a = eng.infer_scenes([engine.Scene(img2, masks2, boxes2), engine.Scene(image, masks, boxes)], "InstaDepthNet_od")
    b = eng.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od")[0]
    c = eng.infer_scenes([engine.Scene(img2, masks2, boxes2)], "InstaDepthNet_od")[0]
    assert np.array_equal(a[1]["occ"], b["occ"]) and np.array_equal(a[1]["depth"], b["depth"])
    assert np.array_equal(a[0]["occ"], c["occ"]) and np.array_equal(a[0]["depth"], c["depth"])
    m = models.InstaDepthNet_od(dict(algo="InstaDepthNet_od", max_pairs=16, max_images=4))
    m.load_state_dict(sd)
    m.switch_to("eval")
    occ, depth = inference.infer_order_sup_occ_depth(m, image, masks, boxes, "all", "InstaDepthNet_od", "resize", G.D, "")
    assert np.array_equal(occ, b["occ"]) and np.array_equal(depth, b["depth"])
    # more images than the encoder handle holds (max_images = 4): batches are cut by image count, results unchanged;
    # images without pairs (one instance) give 1 x 1 zero matrices
    small = [engine.Scene(img2, masks2[k:k + 2], boxes2[k:k + 2]) for k in (0, 1)] * 3 + \
            [engine.Scene(img2, masks2[:1], boxes2[:1])]
    many = eng.infer_scenes(small, "InstaDepthNet_od")
    for k in (0, 1):
        alone = eng.infer_scenes([small[k]], "InstaDepthNet_od")[0]
        for rep in (k, k + 2, k + 4):
            assert np.array_equal(many[rep]["occ"], alone["occ"]) and np.array_equal(many[rep]["depth"], alone["depth"])
    assert many[6]["occ"].shape == (1, 1) and many[6]["occ"][0, 0] == 0


def test_instadepthnet_d_depth_only(setup):
    """InstaDepthNet^d = the same network without oo_net: ``infer_order_sup_depth`` returns the depth matrix of ^od
    (same encoder / do_net weights) and no disparity (disp_select_method '')."""
    z, sd, eng = setup
    image, masks, boxes = G.build_scene()
    ref = eng.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od")[0]
    sd_d = {k: v for k, v in sd.items() if not (k.startswith("module.oo_net") or k.startswith("module.occ_fc"))}
    m = models.InstaDepthNet_d(dict(algo="InstaDepthNet_d", max_pairs=16, max_images=4))
    m.load_state_dict(sd_d)
    m.switch_to("eval")
    depth, disp = inference.infer_order_sup_depth(m, image, masks, boxes, "all", "InstaDepthNet_d", "resize", G.D, "")
    assert disp is None and np.array_equal(depth, ref["depth"])


@pytest.mark.parametrize("sel", ["median", "mean"])
def test_depth_order_from_disparity(golden_dir, sel):
    """disp_select_method = median / mean (reference inference.py:589-599, 79-104) through the reference-shaped API:
    the per-instance depth statistics within 3 % of the reference's (measured 0.8 %; bf16 decoder), the order
    matrix identical wherever the reference's two statistics differ by more than that."""
    z = np.load(os.path.join(golden_dir, "instadepth_disp.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED, with_decoder=True)
    m = models.InstaDepthNet_od(dict(algo="InstaDepthNet_od", max_pairs=16, max_images=2))
    m.load_state_dict(sd)
    m.switch_to("eval")
    image, masks, boxes = G.build_scene()
    order, clipped = inference.infer_order_sup_depth(m, image, masks, boxes, "all", "InstaDepthNet_od", "resize", G.D, sel)
    assert order.shape == (G.N_INST, G.N_INST) and tuple(clipped.shape) == (G.D, G.D)
    _, _, stat = m.engine_for(G.D, disparity=True).disparity_order(engine.Scene(image, masks, boxes), "all", sel)
    ref = z["stat_" + sel]
    rel = np.abs(stat - ref) / ref
    print("depth statistics (%s): max relative deviation %.4f" % (sel, rel.max()))
    assert rel.max() < 3e-2
    checked = 0
    for i in range(G.N_INST):
        for j in range(G.N_INST):
            if i != j and abs(ref[i] - ref[j]) > 0.06 * min(ref[i], ref[j]):
                assert order[i, j] == z["order_" + sel][i, j]
                checked += 1
    # (the fixture's four instances lie within 4 % of each other, so usually no pair qualifies; what the test pins is the
    # statistics themselves and the API contract)
    agree = int((order == z["order_" + sel]).sum())
    print("order matrix entries equal to the reference's: %d / %d, %d beyond the noise margin" % (agree, order.size, checked))


def test_disparity_matches_reference_fixture(golden_dir):
    """The disparity output (encoder layer4 + MiDaS decoder in bf16) against the unmodified reference's fp32 map on the
    fixture image, and pixel-wise against the ideal-bf16 emulation of the same arithmetic."""
    z = np.load(os.path.join(golden_dir, "instadepth_disp.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED, with_decoder=True)
    eng = DepthOrderEngine(G.D, max_pairs=16, max_images=2, with_disparity=True)
    eng.load_state_dict(sd)
    image, masks, boxes = G.build_scene()
    disp = eng.disparity([engine.Scene(image, masks, boxes)])
    assert disp.shape == (1, G.D, G.D) and np.isfinite(disp).all()
    scale = float(z["stats"][1] - z["stats"][0])
    # kernel correctness: against the ideal-bf16 emulation of the same arithmetic, pixel by pixel
    emu = IO.disparity_forward(sd, O.resize_mode_rgb(image, G.D)[None], bf16=True)[0]
    e_px = np.abs(disp[0] - emu).max() / scale
    e_mean = np.abs(disp[0] - emu).mean() / scale
    # precision: against the unmodified reference's fp32 map (bf16 storage through ~45 chained tensors of a random-weight
    # decoder shifts the map by ~4 % of its range -- the emulation shows the same distance)
    pooled, rows = G.disp_digest(disp[0])
    e_p, e_r = np.abs(pooled - z["pooled"]).max() / scale, np.abs(rows - z["rows"]).max() / scale
    print("disparity: |gpu - bf16 emulation| max %.4f mean %.4f of the range; vs fp32 reference: block means %.4f, rows %.4f"
          % (e_px, e_mean, e_p, e_r))
    # two bf16 realisations (tensor-core vs CPU summation order) of a decoder that turns bf16 rounding into a 5 % shift:
    # pixel maximum within that scale, mean difference an order of magnitude below it
    assert e_px < 5e-2 and e_mean < 1e-2
    assert e_p < 8e-2 and e_r < 8e-2
    # the order outputs of the same engine are unchanged by the extra encoder layer
    eng3 = DepthOrderEngine(G.D, max_pairs=16, max_images=2)
    eng3.load_state_dict(sd)
    a = eng.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od")[0]
    b = eng3.infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od")[0]
    assert np.array_equal(a["occ"], b["occ"]) and np.array_equal(a["depth"], b["depth"])


def test_midas_pretrained_ordering(golden_dir):
    """method='midas_pretrained' (reference inference.py:576-583): a plain MiDaS checkpoint (pretrained.* + scratch.*),
    depth order from its disparity map -- with the fixture's encoder / decoder weights it must reproduce the
    InstaDepthNet engine's disparity-based order and the reference's statistics."""
    z = np.load(os.path.join(golden_dir, "instadepth_disp.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED, with_decoder=True)
    sd_midas = {k: v for k, v in sd.items() if k.startswith(("module.pretrained.", "module.scratch."))}
    m = models.MidasNet(dict(algo="midas_pretrained", max_pairs=16, max_images=2))
    m.load_state_dict(sd_midas)
    image, masks, boxes = G.build_scene()
    order, clipped = inference.infer_order_sup_depth(m, image, masks, boxes, "all", "midas_pretrained", "resize", G.D, "median")
    assert np.array_equal(order, z["order_median"])
    with pytest.raises(RuntimeError):
        m.engine_for(G.D, disparity=True).infer_scenes([engine.Scene(image, masks, boxes)], "InstaDepthNet_od")
