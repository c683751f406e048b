"""TEST INFRASTRUCTURE ONLY -- calibrated synthetic checkpoint.

The reference's own random init (xavier, gain 0.02) produces eval-mode logits of ~1e-12, i.e. every decision is an
exact tie, so parity on order matrices needs a checkpoint whose logits are O(1).  ``calibrate`` takes the seeded
kaiming weights of ``instaorder_b200.synth.random_state_dict`` and (1) replaces every BN's running statistics by
the batch statistics measured on synthetic pair crops, (2) rescales the FC heads so logits have std ~2.
The calibrated BN / FC tensors are small and are frozen in ``tests/golden/calib_*.npz`` so that the GPU box
rebuilds *exactly* the same checkpoint (conv weights come from numpy's RandomState, which is portable).
"""
import numpy as np

from instaorder_b200 import synth
from oracle import oracle as O


def calib_inputs(seed=123, n_scenes=2, D=128):
    rng = np.random.RandomState(seed)
    xs = []
    for _ in range(n_scenes):
        image, masks, boxes = synth.make_scene(rng, 240, 320, 4, wh_range=((30, 150), (30, 150)))
        boxes = O.expand_bbox(boxes, 3.0)
        for (i, j) in O.enumerate_pairs(4):
            rgb, mi, mj, _ = O.pair_patch(image, masks, boxes, i, j, D)
            x = O.pair_tensor(rgb, mi, mj)
            xs.append(x)
            xs.append(x[[1, 0, 2, 3, 4]])
    return np.stack(xs).astype(np.float32)


def calibrate(sd, prefix="module.", logit_std=2.0, seed=123):
    """In-place calibration of ``sd`` (numpy arrays).  Returns the dict of tensors that were changed."""
    import torch
    x = calib_inputs(seed)
    changed = {}

    def bn_override(t, name):
        mean = t.mean(dim=(0, 2, 3))
        var = t.var(dim=(0, 2, 3), unbiased=False)
        # keep a spread between batch and running statistics so that BN folding is actually exercised
        rm = (mean * 0.9).numpy().astype(np.float32)
        rv = (var * 1.1 + 1e-3).numpy().astype(np.float32)
        sd[prefix + name + ".running_mean"] = rm
        sd[prefix + name + ".running_var"] = rv
        changed[prefix + name + ".running_mean"] = rm
        changed[prefix + name + ".running_var"] = rv
        w = torch.from_numpy(sd[prefix + name + ".weight"])
        b = torch.from_numpy(sd[prefix + name + ".bias"])
        y = (t - torch.from_numpy(rm)[None, :, None, None]) / torch.sqrt(torch.from_numpy(rv) + 1e-5)[None, :, None, None]
        return y * w[None, :, None, None] + b[None, :, None, None]

    out = O.resnet50_forward(sd, x, prefix=prefix, bn_override=bn_override, return_features=True)
    rng = np.random.RandomState(seed + 1)
    for head in ("fc", "fc_occ", "fc_depth"):
        if head in out:
            s = float(out[head].std()) + 1e-12
            w = (sd[prefix + head + ".weight"] * np.float32(logit_std / s)).astype(np.float32)
            feat_mean = out["features"].mean(axis=0)
            # centre the logits on the calibration set, then add a small random bias
            b = (-(w @ feat_mean) + rng.standard_normal(w.shape[0]) * 0.3).astype(np.float32)
            sd[prefix + head + ".weight"] = w
            sd[prefix + head + ".bias"] = b
            changed[prefix + head + ".weight"] = w
            changed[prefix + head + ".bias"] = b
    return changed


def load_calibrated(npz_path, seed, in_channels=5, num_classes=(2, 3), prefix="module."):
    """Rebuild the calibrated checkpoint from the seed + the frozen calibration tensors."""
    sd = synth.random_state_dict(seed, in_channels, num_classes, prefix)
    with np.load(npz_path) as z:
        for k in z.files:
            assert k in sd and sd[k].shape == z[k].shape, k
            sd[k] = z[k]
    return sd
