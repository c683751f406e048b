"""TEST INFRASTRUCTURE ONLY -- freezes the UNMODIFIED reference's training ``step()`` into tests/golden/train_*.npz.

Run in the build container (needs /root/reference):  ``python -m oracle.gen_golden_train``.

For every case the reference model (``models.__dict__[algo](params)``, reference models/supervised_order.py) is given
the seeded synthetic checkpoint ``instaorder_b200.synth.random_state_dict`` through its own ``load_state``, switched to
train mode, fed the seeded batch of ``oracle.train_oracle.make_batch`` through ``set_input`` and stepped ``n_steps``
times (``step()``: two train-mode forwards, loss, backward, average_gradients, optimiser).  Stored per step: the losses,
and -- as (L2 norm, sum, 64 seeded samples) digests, because the full tensors are 94 MB per step -- every parameter's
gradient, every updated parameter and every BN running statistic; plus the full logits and the full gradients of a
few small tensors.  ``tests/test_train_oracle.py`` pins ``oracle/train_oracle.py`` against these; the ``-m gpu`` tests
compare the CUDA step with the oracle and with these fixtures.
"""
import os
import tempfile

import numpy as np

from instaorder_b200 import synth
from oracle import ref_shim
from oracle import train_oracle as T

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: algo, weight seed, batch seed, B, D, optimiser, steps
    "od_sgd": dict(algo="InstaOrderNet_od", wseed=20, bseed=30, B=4, D=64, optim="SGD", n_steps=2,
                   overlap_weight=0.1, distinct_weight=0.9),
    "od_adam": dict(algo="InstaOrderNet_od", wseed=21, bseed=31, B=3, D=64, optim="Adam", n_steps=2,
                    overlap_weight=1.0, distinct_weight=1.0),
    "d_sgd": dict(algo="InstaOrderNet_d", wseed=22, bseed=32, B=4, D=64, optim="SGD", n_steps=1,
                  overlap_weight=0.3, distinct_weight=0.7),
    "o_sgd": dict(algo="InstaOrderNet_o", wseed=23, bseed=33, B=4, D=64, optim="SGD", n_steps=1),
    "ordernet_sgd": dict(algo="OrderNet", wseed=24, bseed=34, B=4, D=64, optim="SGD", n_steps=1),
    "od_sgd_128": dict(algo="InstaOrderNet_od", wseed=25, bseed=35, B=2, D=128, optim="SGD", n_steps=1,
                       overlap_weight=0.1, distinct_weight=0.9),
    # the shipped ^od training config's input size (384^2: 96 / 48 / 24 / 12-wide feature maps); no reference fixture
    # (used by the teacher-forced emulation test only)
    "od_sgd_384": dict(algo="InstaOrderNet_od", wseed=26, bseed=36, B=2, D=384, optim="SGD", n_steps=1,
                       overlap_weight=0.1, distinct_weight=0.9, golden=False),
}
FULL_GRADS = ["conv1.weight", "bn1.weight", "bn1.bias", "layer1.0.conv1.weight", "layer4.2.bn3.weight"]
LR = 1e-2          # large enough that one update is visible in fp32 digests
WEIGHT_DECAY = 1e-4
BETA1 = 0.9


def case_params(c):
    nc = T.ALGOS[c["algo"]][0]
    return dict(algo=c["algo"], backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=nc),
                optim=c["optim"], lr=LR, weight_decay=WEIGHT_DECAY, beta1=BETA1, use_rgb=True,
                overlap_weight=c.get("overlap_weight", 1.0), distinct_weight=c.get("distinct_weight", 1.0))


def set_input_args(algo, batch):
    if algo == "InstaOrderNet_od":
        return dict(rgb=batch["rgb"], modal1=batch["modal1"], modal2=batch["modal2"], depth_order=batch["depth_order"],
                    count=batch["count"], is_overlap=batch["is_overlap"], occ_order=batch["occ_order"])
    if algo == "InstaOrderNet_d":
        return dict(rgb=batch["rgb"], modal1=batch["modal1"], modal2=batch["modal2"], depth_order=batch["depth_order"],
                    count=batch["count"], is_overlap=batch["is_overlap"])
    return dict(rgb=batch["rgb"], modal1=batch["modal1"], modal2=batch["modal2"], occ_order=batch["occ_order"])


def main():
    import torch
    import torch.distributed as dist
    ns = ref_shim.load()
    if not dist.is_initialized():      # step() always calls utils.average_gradients (all_reduce)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("gloo", rank=0, world_size=1)
    torch.set_num_threads(8)
    for name, c in CASES.items():
        if not c.get("golden", True):
            continue
        algo = c["algo"]
        nc = T.ALGOS[algo][0]
        sd = synth.random_state_dict(c["wseed"], 5, nc)
        model = ns.models.__dict__[algo](case_params(c), dist_model=False)
        with tempfile.TemporaryDirectory() as td:
            torch.save({"step": 0, "state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()},
                        "optimizer": model.optim.state_dict()}, os.path.join(td, "ckpt_iter_0.pth.tar"))
            model.load_state(td, Iter=0)
        model.switch_to("train")
        names = T.param_names(nc)
        plist = dict(model.model.named_parameters())
        out = {}
        for it in range(c["n_steps"]):
            batch = T.make_batch(c["bseed"] + it, c["B"], c["D"], algo)
            logits = []
            hook = model.model.register_forward_hook(
                lambda m, i, o: logits.append(np.concatenate([t.detach().numpy() for t in (o if isinstance(o, tuple) else (o,))], 1)))
            model.set_input(**set_input_args(algo, batch))
            r = model.step()
            hook.remove()
            loss_log, loss = r if isinstance(r, tuple) else ({}, r)
            out["s%d_loss" % it] = np.float32(loss["loss"].item())
            for k, v in loss_log.items():
                out["s%d_%s" % (it, k)] = np.float32(v.item())
            out["s%d_logits" % it] = np.stack(logits).astype(np.float32)          # [2, B, K]
            gn, gs, gsm, pn, ps, psm = [], [], [], [], [], []
            for k in names:
                p = plist["module." + k]
                a, b, s = T.tensor_digest(p.grad.numpy())
                gn.append(a); gs.append(b); gsm.append(s)
                a, b, s = T.tensor_digest(p.detach().numpy())
                pn.append(a); ps.append(b); psm.append(s)
            out["s%d_grad_norm" % it] = np.array(gn)
            out["s%d_grad_sum" % it] = np.array(gs)
            out["s%d_grad_samples" % it] = np.stack(gsm)
            out["s%d_param_norm" % it] = np.array(pn)
            out["s%d_param_sum" % it] = np.array(ps)
            out["s%d_param_samples" % it] = np.stack(psm)
            for k in FULL_GRADS:
                out["s%d_fullgrad_%s" % (it, k)] = plist["module." + k].grad.numpy().astype(np.float32).copy()
            st = model.model.state_dict()
            rs = [k for k in st if k.endswith(("running_mean", "running_var"))]
            out["s%d_stat_norm" % it] = np.array([T.tensor_digest(st[k].numpy())[0] for k in rs])
            out["s%d_stat_samples" % it] = np.stack([T.tensor_digest(st[k].numpy(), 16)[2] for k in rs])
            out["s%d_nbt" % it] = np.int64(st["module.bn1.num_batches_tracked"].item())
        path = os.path.join(GOLDEN, "train_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote", path, "losses", [float(out["s%d_loss" % i]) for i in range(c["n_steps"])],
              os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
