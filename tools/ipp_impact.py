"""One-off measurement (round-1 verdict, weak 3): how much does OpenCV's IPP resize path (cv2's default on x86) change the
UNMODIFIED reference's outputs against the generic path this package is bit-exact with?  Runs the live reference twice on
the golden scenes -- cv2.ipp.setUseIPP(False) / (True) -- and reports changed rgb pixels, logit differences and order-
matrix entries that flip.  Needs /root/reference (build container only).  Output: profiles/r02_ipp_impact.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden as G, ref_shim  # noqa: E402


def run(ns, case, use_ipp):
    import cv2
    cv2.ipp.setUseIPP(use_ipp)
    c = G.CASES[case]
    image, masks, boxes = G.build_scene(case)
    bexp = G.ns_expand(ns, boxes) if c["expand"] else boxes
    model = G.make_reference_model(ns, c["algo"], c["num_classes"], G.state_dict_for(case))
    rec = dict(x=[], logits=[])

    def record(m, inp, out):
        rec["x"].append(inp[0].numpy().copy())
        rec["logits"].append(np.concatenate([o.detach().numpy().reshape(-1) for o in (out if isinstance(out, tuple) else (out,))]))

    hook = model.model.register_forward_hook(record)
    mode, D = c.get("patch_or_image", "patch"), c.get("input_size", 256)
    if c["algo"] == "InstaOrderNet_od":
        occ, depth = ns.inference.infer_order_sup_occ_depth(model, image, masks, bexp, "all", c["algo"], mode, D, "")
    else:
        occ = ns.inference.infer_order_sup_occ(model, image, masks, bexp, "all", c["algo"], mode, D)
        depth = np.zeros_like(occ)
    hook.remove()
    cv2.ipp.setUseIPP(False)
    return np.stack(rec["x"]), np.stack(rec["logits"]), occ, depth


def main():
    ns = ref_shim.load()
    with ref_shim.cpu_only():
        for case in ("c2_od_full", "c2_od_big", "c3_ordernet_full", "c1_o"):
            x0, l0, o0, d0 = run(ns, case, False)
            x1, l1, o1, d1 = run(ns, case, True)
            rgb0, rgb1 = x0[:, 0, 2:], x1[:, 0, 2:]
            changed = float((rgb0 != rgb1).mean())
            print("%-18s pairs %3d | rgb elements that differ %.2f %% (max |diff| %.4f = one u8 step / std) | logits: max |diff| %.5f, "
                  "std %.3f | order entries that flip: occ %d / %d, depth %d / %d" % (
                      case, x0.shape[0] // 2, 100 * changed, float(np.abs(rgb0 - rgb1).max()), float(np.abs(l0 - l1).max()),
                      float(l0.std()), int((o0 != o1).sum()), o0.size, int((d0 != d1).sum()), d0.size))


if __name__ == "__main__":
    main()
