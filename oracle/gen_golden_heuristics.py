"""TEST INFRASTRUCTURE ONLY -- freezes the UNMODIFIED reference's heuristic order baselines (inference.py:272-346:
infer_occ_order_area / _yaxis, infer_depth_order_area / _yaxis, both settings each) on synthetic scenes into
tests/golden/heuristics.npz.  Run in the build container:  ``python -m oracle.gen_golden_heuristics``."""
import os

import numpy as np

from instaorder_b200 import synth
from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SCENES = [dict(seed=3, H=120, W=160, N=6, wh_range=((20, 90), (20, 70))),
          dict(seed=4, H=200, W=150, N=9, wh_range=((15, 80), (15, 120))),
          dict(seed=5, H=64, W=64, N=3, wh_range=((10, 50), (10, 50)))]
VARIANTS = [("infer_occ_order_area", "occluder", ("smaller", "larger")),
            ("infer_occ_order_yaxis", "occluder", ("lower", "higher")),
            ("infer_depth_order_area", "closer", ("smaller", "larger")),
            ("infer_depth_order_yaxis", "closer", ("lower", "higher"))]


def scene_masks(k):
    s = dict(SCENES[k])
    rng = np.random.RandomState(s.pop("seed"))
    return synth.make_scene(rng, **s)[1]


def main():
    infer = ref_shim.load().inference
    out = {}
    for k in range(len(SCENES)):
        masks = scene_masks(k)
        for fn, kw, opts in VARIANTS:
            for o in opts:
                out["s%d_%s_%s" % (k, fn, o)] = np.asarray(getattr(infer, fn)(masks, **{kw: o}), dtype=np.int64)
    path = os.path.join(GOLDEN, "heuristics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "matrices")


if __name__ == "__main__":
    main()
