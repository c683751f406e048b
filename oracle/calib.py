"""TEST INFRASTRUCTURE ONLY -- calibrated synthetic checkpoint.

The reference's own random init (xavier, gain 0.02) produces eval-mode logits of ~1e-12, i.e. every decision is an
exact tie, so parity on order matrices needs a checkpoint whose logits are O(1).  ``calibrate`` takes the seeded
kaiming weights of ``instaorder_b200.synth.random_state_dict`` and (1) replaces every BN's running statistics by
the batch statistics measured on synthetic pair crops, (2) rebuilds the FC heads along the principal directions of the pooled
features with logit std ~0.35 (see the comment in ``calibrate``).
The calibrated BN / FC tensors are small and are frozen in ``tests/golden/calib_*.npz`` so that the GPU box
rebuilds *exactly* the same checkpoint (conv weights come from numpy's RandomState, which is portable).
"""
import numpy as np

from instaorder_b200 import synth
from oracle import oracle as O


def calib_inputs(seed=123, n_scenes=6, D=256, mode="patch"):
    """Synthetic pair tensors (both directions) shaped like the test scenes: COCO-sized images, 4 instances,
    boxes expanded as tools/test.py does."""
    rng = np.random.RandomState(seed)
    xs = []
    for _ in range(n_scenes):
        H, W = synth.COCO_SHAPES[int(rng.randint(0, len(synth.COCO_SHAPES)))]
        image, masks, boxes = synth.make_scene(rng, H, W, 4)
        boxes = O.expand_bbox(boxes, 3.0)
        rgb_whole = O.resize_mode_rgb(image, D) if mode == "resize" else \
            (O.image_mode_rgb(image, D) if mode == "image" else None)
        for (i, j) in O.enumerate_pairs(4):
            if mode == "patch":
                rgb, mi, mj, _ = O.pair_patch(image, masks, boxes, i, j, D)
                x = O.pair_tensor(rgb, mi, mj)
            elif mode == "image":
                x = np.concatenate([O.image_mode_mask(masks[i], D)[None].astype(np.float32),
                                    O.image_mode_mask(masks[j], D)[None].astype(np.float32), rgb_whole])
            else:
                x = np.concatenate([O.resize_mode_mask(masks[i], D)[None].astype(np.float32),
                                    O.resize_mode_mask(masks[j], D)[None].astype(np.float32), rgb_whole])
            xs.append(x)
            xs.append(x[[1, 0, 2, 3, 4]])
    return np.stack(xs).astype(np.float32)


def calibrate(sd, prefix="module.", logit_std=0.35, seed=123, D=256, mode="patch"):
    """In-place calibration of ``sd`` (numpy arrays).  Returns the dict of tensors that were changed."""
    import torch
    x = calib_inputs(seed, D=D, mode=mode)
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
    # Heads: a random deep ReLU net maps all inputs to nearly the same pooled feature (between-sample std is ~16 %
    # of the mean), so a random head scaled to O(1) logits would amplify bf16 rounding noise far more than a trained
    # head does.  Like a trained head, ours reads the directions in which the calibration features actually vary
    # (top principal components), and its scale is kept moderate (logit std ~ logit_std).
    rng = np.random.RandomState(seed + 1)
    feat = out["features"].astype(np.float64)
    mu = feat.mean(axis=0)
    _, _, vt = np.linalg.svd(feat - mu, full_matrices=False)
    for head in ("fc", "fc_occ", "fc_depth"):
        if head in out:
            k = sd[prefix + head + ".weight"].shape[0]
            w = rng.standard_normal((k, 6)) @ vt[:6]
            s = float(((feat - mu) @ w.T).std()) + 1e-12
            w = (w * (logit_std / s)).astype(np.float32)
            b = (-(w.astype(np.float64) @ mu) + rng.standard_normal(k) * 0.1 * logit_std).astype(np.float32)
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
