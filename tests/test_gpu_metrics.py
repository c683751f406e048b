"""Metric kernels vs fixtures produced by the reference's sklearn / numpy code: bit-identical float64."""
import os

import numpy as np
import pytest

from instaorder_b200 import engine, inference, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_metrics_match_reference_fixtures(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    T = len(z["N"])
    cut = lambda a, t: a[t][: z["N"][t], : z["N"][t]]
    for zd in (0, 1):
        idx = [t for t in range(T) if int(z["zd"][t]) == zd]
        got = engine.metrics_prf([cut(z["order"], t) for t in idx], [cut(z["gt"], t) for t in idx], zd)
        assert np.array_equal(got, z["prf"][idx])
    got = engine.metrics_whdr([cut(z["depth_pred"], t) for t in range(T)], [cut(z["gtd"], t) for t in range(T)],
                              [cut(z["ovl"], t) for t in range(T)], [cut(z["cnt"], t) for t in range(T)])
    assert np.array_equal(got, z["whdr"])
    # reference-shaped single-image API
    t = 5
    r = inference.eval_order_recall_precision_f1(cut(z["order"], t), cut(z["gt"], t), int(z["zd"][t]))
    assert tuple(r) == tuple(z["prf"][t])
    w = inference.eval_depth_order_whdr(cut(z["depth_pred"], t), (cut(z["gtd"], t), cut(z["ovl"], t), cut(z["cnt"], t)))
    assert [w[k][0] for k in O.WHDR_KEYS] == list(z["whdr"][t])


def test_metrics_large_batch_vs_oracle():
    """N up to 40 (780 pairs: exercises numpy's recursive pairwise summation) on 300 images."""
    rng = np.random.RandomState(77)
    orders, gts, dps, gds, ovs, cns = [], [], [], [], [], []
    for t in range(300):
        n = int(rng.randint(2, 41))
        occ, depth, ovl, cnt = synth.make_gt(rng, n)
        pred = (rng.rand(n, n) < 0.3).astype(np.int64)
        orders.append(pred); gts.append(occ)
        dps.append(rng.randint(0, 3, size=(n, n)).astype(np.int64)); gds.append(depth); ovs.append(ovl); cns.append(cnt)
    got = engine.metrics_prf(orders, gts, 1)
    want = np.array([O.eval_order_recall_precision_f1(o, g, 1) for o, g in zip(orders, gts)])
    assert np.array_equal(got, want)
    got = engine.metrics_whdr(dps, gds, ovs, cns)
    want = np.array([[float(O.eval_depth_order_whdr(p, (g, v, c))[k][0]) for k in O.WHDR_KEYS]
                     for p, g, v, c in zip(dps, gds, ovs, cns)])
    assert np.array_equal(got, want)


def test_infer_gt_order_matches_reference(golden_dir):
    from oracle import gen_golden
    z = np.load(os.path.join(golden_dir, "gt_order.npz"))
    for t in range(3):
        modal, amodal = gen_golden.kins_scene(int(z["seed%d" % t]))
        got = inference.infer_gt_order(modal, amodal)
        assert got.dtype == np.int64 and np.array_equal(got, z["gt%d" % t])


def test_heuristic_baselines_match_reference(golden_dir):
    """infer_occ_order_area / _yaxis and infer_depth_order_area / _yaxis (reference inference.py:272-346) through the
    io_mask_stats + io_pair_bordering kernels: matrices identical to the unmodified reference's (integer work)."""
    import os
    import numpy as np
    from instaorder_b200 import inference as infer
    from oracle import gen_golden_heuristics as GH
    z = np.load(os.path.join(golden_dir, "heuristics.npz"))
    for k in range(len(GH.SCENES)):
        masks = GH.scene_masks(k)
        for fn, kw, opts in GH.VARIANTS:
            for o in opts:
                got = getattr(infer, fn)(masks, **{kw: o})
                assert np.array_equal(got, z["s%d_%s_%s" % (k, fn, o)]), (k, fn, o)
