"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of the reference's *training step* for the order networks.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file; the product path
(``instaorder_b200/``) never does.

Follows (paths relative to /root/reference):
  * ``models/supervised_order.py:32-48, 383-395, 451-463, 509-516``  set_input (swapped-direction labels)
  * ``models/supervised_order.py:83-95, 413-438, 481-493, 535-548``  step(): two train-mode forwards, loss,
    zero_grad / backward / average_gradients / optim.step
  * ``models/supervised_order.py:60-81``                             calculate_loss (softmax -> CrossEntropyLoss,
    sigmoid -> BCELoss, overlap / distinct weighting, / world_size)
  * ``models/backbone/resnet_cls.py:75-222``                         the 5-channel ResNet-50 in train mode
  * ``models/single_stage_model.py:34-42``                           SGD(momentum 0.9, weight_decay) / Adam(beta1, 0.999)

Pinning: ``oracle/gen_golden_train.py`` ran the UNMODIFIED reference ``InstaOrderNet_od / _d / _o / OrderNet .step()``
(through ``oracle/ref_shim.py``) in the build container on seeded inputs and froze losses, gradient / parameter
digests and running statistics into ``tests/golden/train_*.npz``; ``tests/test_train_oracle.py`` checks this
restatement against them.
"""
import collections

import numpy as np

from instaorder_b200 import synth

ALGOS = {
    # algo: (num_classes, occ head name, class head name, uses overlap masks in step())
    "InstaOrderNet_od": ([2, 3], "fc_occ", "fc_depth", True),
    "InstaOrderNet_d": (3, None, "fc", True),
    "InstaOrderNet_o": (2, "fc", None, False),
    "OrderNet": (3, None, "fc", False),
}


def param_names(num_classes, in_channels=5):
    """Names of the trainable tensors in ``model.parameters()`` order (running stats excluded)."""
    return [k for k, _ in synth.resnet50_layout(in_channels, num_classes)
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


def make_batch(seed, B, D, algo):
    """Seeded synthetic training batch in the ``set_input`` argument types."""
    import torch
    rng = np.random.RandomState(seed)
    rgb = rng.standard_normal((B, 3, D, D)).astype(np.float32)
    masks = []
    yy, xx = np.mgrid[0:D, 0:D]
    for _ in range(2 * B):
        cx, cy = rng.uniform(0.2, 0.8, 2) * D
        rx, ry = rng.uniform(0.1, 0.4, 2) * D
        masks.append((((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0).astype(np.float32))
    masks = np.stack(masks).reshape(2, B, 1, D, D)
    batch = dict(rgb=torch.from_numpy(rgb), modal1=torch.from_numpy(masks[0]), modal2=torch.from_numpy(masks[1]))
    if algo in ("InstaOrderNet_od", "InstaOrderNet_d"):
        batch["depth_order"] = torch.from_numpy(rng.randint(0, 3, B).astype(np.int64))
        batch["count"] = torch.from_numpy(rng.randint(2, 4, B).astype(np.int64))
        ov = (rng.rand(B) < 0.3).astype(np.int64)
        if B >= 2:
            ov[0], ov[1] = 1, 0          # both subsets non-empty
        batch["is_overlap"] = torch.from_numpy(ov)
    if algo in ("InstaOrderNet_od", "InstaOrderNet_o"):
        batch["occ_order"] = torch.from_numpy((rng.rand(B, 2) < 0.3).astype(np.float32))
    if algo == "OrderNet":
        batch["occ_order"] = torch.from_numpy(rng.randint(0, 3, B).astype(np.int64))
    return batch


def swap01(t):
    """order2 of set_input: 0 -> 1, 1 -> 0, everything else unchanged (supervised_order.py:40-41, 391-392, 460-463)."""
    o = t.clone()
    o[t == 0] = 1
    o[t == 1] = 0
    return o


def forward_train(P, S, x, eps=1e-5, momentum=0.1):
    """One train-mode forward (resnet_cls.py:203-222).  P: name -> parameter tensor (requires_grad), S: name -> running
    stat tensor (updated in place, like nn.BatchNorm2d).  Returns dict head -> logits."""
    import torch
    import torch.nn.functional as F

    def bn(t, name):
        return F.batch_norm(t, S[name + ".running_mean"], S[name + ".running_var"], P[name + ".weight"],
                            P[name + ".bias"], True, momentum, eps)

    t = F.relu(bn(F.conv2d(x, P["conv1.weight"], stride=2, padding=3), "bn1"))
    t = F.max_pool2d(t, 3, 2, 1)
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            p = "layer%d.%d" % (li, b)
            stride = 2 if (b == 0 and li > 1) else 1
            idt = t
            o = F.relu(bn(F.conv2d(t, P[p + ".conv1.weight"]), p + ".bn1"))
            o = F.relu(bn(F.conv2d(o, P[p + ".conv2.weight"], stride=stride, padding=1), p + ".bn2"))
            o = bn(F.conv2d(o, P[p + ".conv3.weight"]), p + ".bn3")
            if b == 0:
                idt = bn(F.conv2d(t, P[p + ".downsample.0.weight"], stride=stride), p + ".downsample.1")
            t = F.relu(o + idt)
    feat = torch.flatten(F.adaptive_avg_pool2d(t, 1), 1)
    out = {}
    for head in ("fc", "fc_occ", "fc_depth"):
        if head + ".weight" in P:
            out[head] = F.linear(feat, P[head + ".weight"], P[head + ".bias"])
    return out


def step_loss(algo, out1, out2, batch, overlap_weight=1.0, distinct_weight=1.0, world_size=1):
    """The loss of ``step()`` for each class; returns (loss, occ_loss, class_loss) tensors."""
    import torch
    import torch.nn.functional as F
    _, occ_head, cls_head, use_masks = ALGOS[algo]
    zero = torch.zeros(())
    occ_loss, cls_loss = zero, zero
    if occ_head is not None:
        o1, o2 = torch.sigmoid(out1[occ_head]), torch.sigmoid(out2[occ_head])      # :86 / :537
        t1 = batch["occ_order"]
        t2 = t1[:, [1, 0]]                                                          # :47-48 / :516
        occ_loss = F.binary_cross_entropy(o1, t1) + F.binary_cross_entropy(o2, t2)  # :75-76 / :543
    if cls_head is not None:
        p1, p2 = F.softmax(out1[cls_head], dim=1), F.softmax(out2[cls_head], dim=1)  # :85 / :415 / :483
        y1 = batch["depth_order"] if algo != "OrderNet" else batch["occ_order"]
        y2 = swap01(y1)
        if use_masks:                                                                # :62-73 / :421-433
            ovl, dis = batch["is_overlap"] == 1, batch["is_overlap"] == 0
            lo = F.cross_entropy(p1[ovl], y1[ovl]) + F.cross_entropy(p2[ovl], y2[ovl]) if ovl.sum() > 0 else zero
            ld = F.cross_entropy(p1[dis], y1[dis]) + F.cross_entropy(p2[dis], y2[dis]) if dis.sum() > 0 else zero
            cls_loss = lo * overlap_weight + ld * distinct_weight
        else:                                                                        # :488
            cls_loss = F.cross_entropy(p1, y1) + F.cross_entropy(p2, y2)
    return (cls_loss + occ_loss) / world_size, occ_loss, cls_loss


def train_step(sd, batch, algo, lr=1e-4, weight_decay=1e-4, optim="SGD", beta1=0.9, opt_state=None,
               overlap_weight=1.0, distinct_weight=1.0, world_size=1, prefix="module.", apply_update=True):
    """One ``step()``.  sd: reference-layout state_dict (numpy / torch); returns a dict with the losses, the gradients
    (name -> numpy, reference layout), the updated state_dict (numpy) and the optimiser state."""
    import torch
    num_classes = ALGOS[algo][0]
    names = param_names(num_classes)

    def as_t(v):
        return v.clone() if isinstance(v, torch.Tensor) else torch.from_numpy(np.array(v))

    P = {k: as_t(sd[prefix + k]).float().requires_grad_(True) for k in names}
    S = {k[len(prefix):]: as_t(v).float() for k, v in sd.items()
         if k.endswith(("running_mean", "running_var"))}
    x1 = torch.cat([batch["modal1"], batch["modal2"], batch["rgb"]], dim=1)          # :84
    x2 = torch.cat([batch["modal2"], batch["modal1"], batch["rgb"]], dim=1)
    out1 = forward_train(P, S, x1)
    out2 = forward_train(P, S, x2)
    loss, occ_loss, cls_loss = step_loss(algo, out1, out2, batch, overlap_weight, distinct_weight, world_size)
    loss.backward()
    grads = {k: P[k].grad.detach().numpy().copy() for k in names}
    res = dict(loss=float(loss.detach()), loss_occ=float(occ_loss.detach()), loss_cls=float(cls_loss.detach()), grads=grads,
               logits1={k: v.detach().numpy() for k, v in out1.items()},
               logits2={k: v.detach().numpy() for k, v in out2.items()})
    new_sd = collections.OrderedDict()
    state = opt_state if opt_state is not None else {"step": 0, "bufs": {}}
    if apply_update:
        state["step"] += 1
        with torch.no_grad():
            for k in names:
                w, g = P[k].detach(), P[k].grad
                if optim == "SGD":                      # torch.optim.SGD, momentum 0.9 (single_stage_model.py:34-38)
                    g = g + weight_decay * w
                    buf = state["bufs"].get(k)
                    buf = g.clone() if buf is None else 0.9 * buf + g
                    state["bufs"][k] = buf
                    w = w - lr * buf
                else:                                   # torch.optim.Adam(betas=(beta1, 0.999)) (:39-42)
                    m, v = state["bufs"].get(k, (torch.zeros_like(w), torch.zeros_like(w)))
                    m = beta1 * m + (1 - beta1) * g
                    v = 0.999 * v + 0.001 * g * g
                    state["bufs"][k] = (m, v)
                    t = state["step"]
                    denom = v.sqrt() / np.sqrt(1 - 0.999 ** t) + 1e-8
                    w = w - (lr / (1 - beta1 ** t)) * (m / denom)
                new_sd[prefix + k] = w.numpy().copy()
    else:
        for k in names:
            new_sd[prefix + k] = P[k].detach().numpy().copy()
    for k, v in S.items():
        new_sd[prefix + k] = v.numpy().copy()
    res["state_dict"] = new_sd
    res["opt_state"] = state
    return res


def tensor_digest(a, n_samples=64, seed=0):
    """Small fingerprint of a tensor for the golden fixtures: (L2 norm, sum, n_samples seeded elements)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.random.RandomState(seed).randint(0, a.size, size=n_samples)
    return np.float64(np.sqrt((a * a).sum())), np.float64(a.sum()), a[idx].astype(np.float32)
