"""The CPU restatement of InstaDepthNet^od's order branch (oracle/instadepth_oracle.py) against the fixture frozen from
the UNMODIFIED reference (oracle/gen_golden_instadepth.py: reference ``infer_order_sup_occ_depth`` with
method='InstaDepthNet_od', resize 384^2): logits of all 12 forwards and the two order matrices."""
import os

import numpy as np

from oracle import gen_golden_instadepth as G, instadepth_oracle as IO, oracle as O


def reference_inputs():
    image, masks, boxes = G.build_scene()
    rgb = O.resize_mode_rgb(image, G.D)[None]
    mm = [O.resize_mode_mask(m, G.D)[None].astype(np.float32) for m in masks]
    m1, m2 = [], []
    for (i, j) in O.enumerate_pairs(G.N_INST):
        m1 += [mm[i], mm[j]]
        m2 += [mm[j], mm[i]]
    return image, masks, boxes, rgb, np.stack(m1), np.stack(m2)


def matrices_from_logits(dl, ol, n):
    """reference inference.py:107-137 (probabilities averaged over the two directions) and :416-434 (matrix writes)"""
    occ = np.zeros((n, n), np.int64)
    depth = np.zeros((n, n), np.int64)
    for p, (i, j) in enumerate(O.enumerate_pairs(n)):
        arg, _ = O.decide_depth(dl[p, 0], dl[p, 1])
        a_over_b, b_over_a, _ = O.decide_occ(ol[p, 0], ol[p, 1])
        O.write_depth(depth, i, j, arg)
        O.write_occ(occ, i, j, a_over_b, b_over_a)
    return occ, depth


def test_oracle_matches_reference_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "instadepth_order.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED)
    _, _, _, rgb, m1, m2 = reference_inputs()
    P = G.N_INST * (G.N_INST - 1) // 2
    out = IO.order_forward(sd, rgb, m1, m2, np.zeros(2 * P, np.int64))
    dl, ol = out["depth"].reshape(P, 2, 3), out["occ"].reshape(P, 2, 2)
    assert np.abs(dl - z["depth_logits"]).max() < 2e-4
    assert np.abs(ol - z["occ_logits"]).max() < 2e-4
    occ, depth = matrices_from_logits(dl, ol, G.N_INST)
    assert np.array_equal(occ, z["occ"]) and np.array_equal(depth, z["depth"])
    assert z["depth_logits"].std() > 0.05 and z["occ_logits"].std() > 0.05      # not the all-tie random init


def test_disparity_oracle_matches_reference_fixture(golden_dir):
    """The disparity branch (encoder layer4 + MiDaS decoder) of the restatement against the reference's output frozen
    in instadepth_disp.npz -- groundwork: the CUDA path does not compute the disparity yet."""
    z = np.load(os.path.join(golden_dir, "instadepth_disp.npz"))
    sd = IO.load_calibrated(os.path.join(golden_dir, "instadepth_calib.npz"), G.SEED, with_decoder=True)
    image, _, _ = G.build_scene()
    disp = IO.disparity_forward(sd, O.resize_mode_rgb(image, G.D)[None])[0]
    pooled, rows = G.disp_digest(disp)
    scale = float(z["stats"][1] - z["stats"][0])
    assert np.abs(pooled - z["pooled"]).max() < 1e-4 * scale
    assert np.abs(rows - z["rows"]).max() < 1e-4 * scale
    assert disp.min() >= 0.0 and z["stats"][3] > 0.1        # non_negative=True output with real structure
