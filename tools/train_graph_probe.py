"""Probe: does replaying the training step as a CUDA graph (torch.cuda.CUDAGraph capture of the library's launches) cut
the per-launch gaps?  Usage: python tools/train_graph_probe.py [batch] [size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import synth, training  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
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


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(3):
    step()          # first_step flag of SGD flips after the first call: capture a steady-state step
torch.cuda.synchronize()
print("eager : %.3f ms / step, loss %.4f" % (timeit(step), float(eng.losses[0])))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
graph = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=s):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print("graph : %.3f ms / step, loss %.4f" % (timeit(graph.replay), float(eng.losses[0])))
