"""Where does the end-to-end time of OrderEngine.infer_scenes go?  Host timers around the phases of one 17-image call
(765 pairs) of the bench workload, GPU time from CUDA events, averaged over a few calls."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import engine, synth  # noqa: E402

T = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


def main():
    n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    scenes = []
    for image, masks, boxes in synth.coco_scene_stream(1000, 17 * 4, N=10):
        scenes.append(engine.Scene(image, masks, engine.expand_bbox(boxes, 3.0)))
    eng = engine.OrderEngine([2, 3], 256, max_pairs=256)
    eng.load_state_dict(synth.random_state_dict(0, 5, [2, 3]))
    for i in range(3):
        eng.infer_scenes(scenes[:17], "InstaOrderNet_od")
    torch.cuda.synchronize()
    eng.stage_batch = timed("stage_batch (pack into pinned + enqueue H2D)", eng.stage_batch)
    eng.gather = timed("gather (enqueue)", eng.gather)
    eng.forward = timed("forward (enqueue 39 launches)", eng.forward)
    eng.decide = timed("decide (enqueue)", eng.decide)
    engine.enumerate_pairs = timed("enumerate_pairs", engine.enumerate_pairs)
    engine.pair_crop_boxes = timed("pair_crop_boxes", engine.pair_crop_boxes)
    t0 = time.perf_counter()
    for c in range(n_calls):
        sc = [scenes[(17 * c + k) % len(scenes)] for k in range(17)]
        eng.infer_scenes(sc, "InstaOrderNet_od")
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print("calls %d, %.2f ms per call of 765 pairs -> %.0f pairs/s" % (n_calls, 1000 * total / n_calls, 765 * n_calls / total))
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
        print("  %-50s %.3f ms per call" % (k, 1000 * v / n_calls))
    print("  %-50s %.3f ms per call" % ("everything else (python, final D2H + sync)", 1000 * (total - sum(T.values())) / n_calls))


if __name__ == "__main__":
    main()
