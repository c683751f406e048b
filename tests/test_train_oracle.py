"""Pins the training-step oracle (oracle/train_oracle.py, torch-CPU fp32) against the fixtures produced by the
UNMODIFIED reference ``step()`` (oracle/gen_golden_train.py -> tests/golden/train_*.npz): losses, every gradient,
every updated parameter and every BN running statistic (digests), the logits and a few full gradient tensors.
Floating point, same torch ops in a different composition -> tolerance 2e-4 relative to each tensor's scale."""
import os

import numpy as np
import pytest

from instaorder_b200 import synth
from oracle import gen_golden_train as G
from oracle import train_oracle as T

RTOL = 2e-4


def _close(got, want, scale, what):
    err = float(np.max(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64))))
    assert err <= RTOL * scale + 1e-7, "%s: err %.4g (scale %.4g)" % (what, err, scale)


@pytest.mark.parametrize("case", ["od_sgd", "od_adam", "d_sgd", "o_sgd", "ordernet_sgd"])
def test_oracle_step_matches_reference(case, golden_dir):
    c = G.CASES[case]
    algo = c["algo"]
    nc = T.ALGOS[algo][0]
    z = np.load(os.path.join(golden_dir, "train_%s.npz" % case))
    sd = synth.random_state_dict(c["wseed"], 5, nc)
    names = T.param_names(nc)
    state = None
    for it in range(c["n_steps"]):
        batch = T.make_batch(c["bseed"] + it, c["B"], c["D"], algo)
        r = T.train_step(sd, batch, algo, lr=G.LR, weight_decay=G.WEIGHT_DECAY, optim=c["optim"], beta1=G.BETA1,
                         opt_state=state, overlap_weight=c.get("overlap_weight", 1.0),
                         distinct_weight=c.get("distinct_weight", 1.0))
        state = r["opt_state"]
        pre = "s%d_" % it
        _close(r["loss"], z[pre + "loss"], abs(float(z[pre + "loss"])), "loss")
        if algo == "InstaOrderNet_od":
            _close(r["loss_occ"], z[pre + "loss_occ"], 1.0, "loss_occ")
            _close(r["loss_cls"], z[pre + "loss_depth"], 1.0, "loss_depth")
        heads = [h for h in ("fc", "fc_occ", "fc_depth") if h in r["logits1"]]
        lg = np.stack([np.concatenate([r["logits1"][h] for h in heads], 1),
                       np.concatenate([r["logits2"][h] for h in heads], 1)])
        _close(lg, z[pre + "logits"], float(np.abs(z[pre + "logits"]).max()), "logits")
        for i, k in enumerate(names):
            gn, _, gs = T.tensor_digest(r["grads"][k])
            gscale = float(z[pre + "grad_norm"][i]) / np.sqrt(r["grads"][k].size) * 10 + 1e-12
            _close(gn, z[pre + "grad_norm"][i], float(z[pre + "grad_norm"][i]) * 5, "grad norm " + k)
            _close(gs, z[pre + "grad_samples"][i], gscale * 5, "grad samples " + k)
            pn, _, ps = T.tensor_digest(r["state_dict"]["module." + k])
            _close(pn, z[pre + "param_norm"][i], float(z[pre + "param_norm"][i]), "param norm " + k)
            _close(ps, z[pre + "param_samples"][i], float(np.abs(z[pre + "param_samples"][i]).max()) + 1e-3,
                   "param samples " + k)
        for k in G.FULL_GRADS:
            want = z[pre + "fullgrad_" + k]
            _close(r["grads"][k], want, float(np.abs(want).max()) * 5, "full grad " + k)
        rs = [k for k in r["state_dict"] if k.endswith(("running_mean", "running_var"))]
        got = np.array([T.tensor_digest(r["state_dict"][k])[0] for k in rs])
        _close(got, z[pre + "stat_norm"], float(z[pre + "stat_norm"].max()), "running stats")
        sd = r["state_dict"]
