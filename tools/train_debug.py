"""Layer-by-layer comparison of the CUDA training forward with the bf16-storage torch emulation (bring-up tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from instaorder_b200 import synth, training  # noqa: E402
from oracle import gen_golden_train as G  # noqa: E402
from oracle import train_oracle as T  # noqa: E402
import gpu_util as U  # noqa: E402

DEV = "cuda:0"
case = sys.argv[1] if len(sys.argv) > 1 else "od_sgd"
c = G.CASES[case]
algo = c["algo"]
nc = T.ALGOS[algo][0]
sd = synth.random_state_dict(c["wseed"], 5, nc)
B, D = c["B"], c["D"]
eng = training.TrainEngine(nc, D, B, DEV)
eng.load_state_dict(sd)
batch = T.make_batch(c["bseed"], B, D, algo)
eng.pack_inputs(batch["rgb"], batch["modal1"], batch["modal2"])
heads = {"InstaOrderNet_od": (0, 2, 3), "InstaOrderNet_d": (-1, 0, 3), "InstaOrderNet_o": (0, -1, 0), "OrderNet": (-1, 0, 3)}[algo]
occ = batch["occ_order"].to(DEV).float().contiguous() if algo in ("InstaOrderNet_od", "InstaOrderNet_o") else None
cls = batch["occ_order"].to(DEV).contiguous() if algo == "OrderNet" else (
    batch["depth_order"].to(DEV).contiguous() if "depth_order" in batch else None)
ovl = batch["is_overlap"].to(DEV).contiguous() if T.ALGOS[algo][3] else None
eng.forward_backward(heads[0], heads[1], heads[2], occ, cls, ovl, 1.0, 1.0, 1, backward=False)
torch.cuda.synchronize()

names = T.param_names(nc)
P = {k: torch.from_numpy(np.asarray(sd["module." + k])).to(DEV) for k in names}
S = {k[7:]: torch.from_numpy(np.asarray(v)).to(DEV).clone() for k, v in sd.items() if k.endswith(("running_mean", "running_var"))}
bt = {k: v.to(DEV) for k, v in batch.items()}
recs = []
with torch.no_grad():
    for x in (torch.cat([bt["modal1"], bt["modal2"], bt["rgb"]], 1), torch.cat([bt["modal2"], bt["modal1"], bt["rgb"]], 1)):
        r = {}
        U.emulated_forward_train(P, S, x, record=r)
        recs.append(r)
for key in recs[0]:
    name, which = key.rsplit(".", 1)
    want = torch.cat([recs[0][key], recs[1][key]], 0).permute(0, 2, 3, 1).contiguous()   # [2B, h, w, c]
    got = eng.activation(name, 0 if which == "y" else 1).float().view(want.shape)
    err = (got - want).abs()
    print("%-28s %s  shape %-22s max|ref| %8.4f  max err %8.4f  mean err %.5f  frac>1e-2*(1+|ref|) %.4f" % (
        name, which, tuple(want.shape), float(want.abs().max()), float(err.max()), float(err.mean()),
        float((err > 1e-2 * (1 + want.abs())).float().mean())))
