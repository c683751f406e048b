"""Per-launch report of one forward pass (CUDA events bracketed around every kernel by io_net_profile):
ms, TFLOP/s and GB/s against the measured peaks.  Usage: python tools/layer_report.py [pairs] > report.txt"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import _lib, engine, synth  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(bf16_tflops_sustained=1400.0, hbm_gbs=6650.0)
    eng = engine.OrderEngine([2, 3], 256, max_pairs=P)
    eng.load_state_dict(synth.random_state_dict(0, 5, [2, 3]))
    eng.pair_tensor.random_(0, 64)       # any finite bf16 pattern will do for timing
    _lib.check(eng.lib.io_net_profile(eng.net, 1))
    for _ in range(3):
        eng.forward(P)
    torch.cuda.synchronize()
    mx = 8192
    ms = np.zeros(mx, np.float32); kind = np.zeros(mx, np.int32); fl = np.zeros(mx, np.float64)
    by = np.zeros(mx, np.float64); tag = np.zeros(mx, np.int32)
    n = _lib.check(eng.lib.io_net_profile_read(eng.net, _lib.ptr(ms), _lib.ptr(kind), _lib.ptr(fl), _lib.ptr(by),
                                               _lib.ptr(tag), mx))
    agg = {}
    for i in range(n):
        a = agg.setdefault(int(tag[i]), [0.0, 0.0, 0.0, 0])
        a[0] += ms[i]; a[1] += fl[i]; a[2] += by[i]; a[3] += 1
    tot = sum(a[0] for a in agg.values())
    print("pairs %d, launches %d, sum of kernel times %.3f ms -> %.0f pairs/s if back-to-back" % (P, n, tot, P / tot * 1e3))
    print("%6s %5s %9s %9s %9s %7s %7s" % ("tag", "n", "ms", "TFLOP/s", "GB/s", "%tc", "%hbm"))
    for t in sorted(agg):
        m, f, b, c = agg[t]
        print("%6d %5d %9.4f %9.1f %9.1f %7.1f %7.1f" % (t, c, m, f / m / 1e9, b / m / 1e6,
              100 * f / m / 1e9 / peaks["bf16_tflops_sustained"], 100 * b / m / 1e6 / peaks["hbm_gbs"]))


if __name__ == "__main__":
    main()
