"""TEST INFRASTRUCTURE ONLY -- tests/golden/order_c2_od_orig.npz: the UNMODIFIED reference's
``infer_order_sup_occ_depth(..., patch_or_image="orig")`` (inference.py:401-408: the image is resized to the nearest
multiples of 32 of its own H and W, a NON-SQUARE network input) on a synthetic scene, with the calibrated checkpoint of
the ``c2_od_resize`` case.  Groundwork for SURVEY.md row G12: the CUDA path does not run ``orig`` yet;
the oracle restatement (``oracle.infer_order(..., "orig")``) is pinned by tests/test_oracle_golden.py.

    python -m oracle.gen_golden_orig"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import calib, gen_golden, ref_shim  # noqa: E402

CASE = "c2_od_resize"       # weights / scene of this case; 333 x 500 -> network input 320 x 512


def main():
    import torch
    ns = ref_shim.load()
    c = gen_golden.CASES[CASE]
    sd = calib.load_calibrated(gen_golden.calib_path(CASE), c["wseed"], 5, c["num_classes"])
    model = gen_golden.make_reference_model(ns, c["algo"], c["num_classes"], sd)
    image, masks, boxes = gen_golden.build_scene(CASE)
    calls = []
    inner = model.model

    class Rec(torch.nn.Module):
        def forward(self, x):
            out = inner(x)
            calls.append((tuple(x.shape), out[0].numpy().copy(), out[1].numpy().copy()))
            return out

    model.model = Rec()
    occ, depth = ns.inference.infer_order_sup_occ_depth(model, image, masks, boxes, "all", "InstaOrderNet_od", "orig",
                                                        c["input_size"], "")
    P = masks.shape[0] * (masks.shape[0] - 1) // 2
    assert len(calls) == 2 * P
    shape = calls[0][0]
    l0 = np.stack([k[1][0] for k in calls]).reshape(P, 2, 2)
    l1 = np.stack([k[2][0] for k in calls]).reshape(P, 2, 3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "order_c2_od_orig.npz"), occ=occ.astype(np.int64),
                        depth=depth.astype(np.int64), logits0=l0.astype(np.float32), logits1=l1.astype(np.float32),
                        net_input_shape=np.asarray(shape, np.int64))
    print("network input", shape, "\nocc\n", occ, "\ndepth\n", depth)


if __name__ == "__main__":
    main()
