"""Per-parameter gradient error of the CUDA training step against the teacher-forced bf16 emulation (bring-up tool)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train_step as TS  # noqa: E402
from oracle import gen_golden_train as G  # noqa: E402
from oracle import train_oracle as T  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "od_sgd"
c = G.CASES[case]
nc = T.ALGOS[c["algo"]][0]
eng, sd, batch, losses = TS._run_engine(c)
loss, logits, grads_ref, S, dev_log = TS._emulate(c, sd, batch, eng)
grads = eng.export_flat(eng.grads, params_only=True)
for k in T.param_names(nc):
    g, w = grads[k].to(TS.DEV), grads_ref[k]
    rel = float((g - w).norm() / (w.norm() + 1e-12))
    cos = float((g * w).sum() / (g.norm() * w.norm() + 1e-20))
    print("%-34s |ref| %10.4g  rel-L2 err %8.5f  1-cos %9.2e  |got|/|ref| %.4f" % (
        k, float(w.norm()), rel, 1 - cos, float(g.norm() / (w.norm() + 1e-20))))
