"""Per-launch report of one training step (CUDA events around every launch, io_train_profile): ms, TFLOP/s and GB/s
against the measured peaks, by kind (conv forward / data gradient / weight gradient / element-wise) and by layer.
Usage: python tools/train_report.py [batch_pairs] [input_size] [steps] > report.txt"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import _lib, synth, training  # noqa: E402

KINDS = {0: "conv fwd", 1: "dgrad", 2: "wgrad", 3: "elementwise", 4: "loss"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(bf16_tflops_sustained=1400.0, hbm_gbs=6650.0)
    eng = training.TrainEngine([2, 3], D, B)
    eng.load_state_dict(synth.random_state_dict(0, 5, [2, 3]))
    opt = training.FlatOptim("SGD", 1e-4, weight_decay=1e-4)
    opt.attach(eng)
    g = torch.Generator(device="cuda").manual_seed(0)
    rgb = torch.randn((B, 3, D, D), generator=g, device="cuda")
    m1 = (torch.rand((B, 1, D, D), generator=g, device="cuda") > 0.7).float()
    m2 = (torch.rand((B, 1, D, D), generator=g, device="cuda") > 0.7).float()
    occ = (torch.rand((B, 2), generator=g, device="cuda") < 0.2).float()
    dep = torch.randint(0, 3, (B,), generator=g, device="cuda")
    ovl = (torch.rand((B,), generator=g, device="cuda") < 0.3).long()

    def step():
        eng.pack_inputs(rgb, m1, m2)
        eng.forward_backward(0, 2, 3, occ, dep, ovl, 0.1, 0.9, 1)
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    flops_step = 3 * 21.764e9 * B * (D / 256.0) ** 2
    print("batch %d pairs @ %d^2: %.3f ms / step -> %.0f pairs/s, %.1f TFLOP/s algorithmic (3 x forward FLOPs), "
          "loss %.4f, %d launches" % (B, D, ms_step, B / ms_step * 1e3, flops_step / ms_step / 1e9,
                                      float(eng.losses[0]), eng.lib.io_train_last_launches(eng.handle)))
    _lib.check(eng.lib.io_train_profile(eng.handle, 1))
    step()
    torch.cuda.synchronize()
    mx = 4096
    ms = np.zeros(mx, np.float32); kind = np.zeros(mx, np.int32); fl = np.zeros(mx, np.float64)
    by = np.zeros(mx, np.float64); tag = np.zeros(mx, np.int32)
    n = _lib.check(eng.lib.io_train_profile_read(eng.handle, _lib.ptr(ms), _lib.ptr(kind), _lib.ptr(fl), _lib.ptr(by),
                                                 _lib.ptr(tag), mx))
    tot = float(ms[:n].sum())
    print("profiled step: %d ops, sum of op times %.3f ms" % (n, tot))
    print("%-12s %5s %9s %7s %9s %9s %7s %7s" % ("kind", "n", "ms", "share", "TFLOP/s", "GB/s", "%tc", "%hbm"))
    for k in sorted(KINDS):
        sel = kind[:n] == k
        if not sel.any():
            continue
        m, f, b = float(ms[:n][sel].sum()), float(fl[:n][sel].sum()), float(by[:n][sel].sum())
        print("%-12s %5d %9.3f %6.1f%% %9.1f %9.1f %7.1f %7.1f" % (
            KINDS[k], int(sel.sum()), m, 100 * m / tot, f / m / 1e9, b / m / 1e6,
            100 * f / m / 1e9 / peaks["bf16_tflops_sustained"], 100 * b / m / 1e6 / peaks["hbm_gbs"]))
    print("\n%6s %-12s %4s %9s %9s %9s %7s %7s" % ("tag", "kind", "n", "ms", "TFLOP/s", "GB/s", "%tc", "%hbm"))
    agg = {}
    for i in range(n):
        a = agg.setdefault((int(tag[i]), int(kind[i])), [0.0, 0.0, 0.0, 0])
        a[0] += ms[i]; a[1] += fl[i]; a[2] += by[i]; a[3] += 1
    for (t, k) in sorted(agg):
        m, f, b, c = agg[(t, k)]
        print("%6d %-12s %4d %9.4f %9.1f %9.1f %7.1f %7.1f" % (
            t, KINDS[k], c, m, f / m / 1e9, b / m / 1e6, 100 * f / m / 1e9 / peaks["bf16_tflops_sustained"],
            100 * b / m / 1e6 / peaks["hbm_gbs"]))


if __name__ == "__main__":
    main()
