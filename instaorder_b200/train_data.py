"""Training-side pair construction on the GPU.

The reference builds every training sample in a DataLoader worker (``datasets/depth_occ_order_dataset.py:142-252``,
``depth_order_dataset.py``, ``occ_order_dataset.py``): pick the pair, jitter / rescale the union crop, crop + cv2 resize
the image and both masks, flip, swap A/B, normalise -- per sample, on the CPU.  Here the host only draws the random
numbers (same ``np.random`` order as the reference, so a run is reproducible against it) and the label logic; the crop
/ bicubic + nearest resize / flip / normalise / concatenation of a whole batch is ONE launch of the fused gather
kernel (``io_pair_gather_patch``, flip flag in ``io_pair_desc.rgb_slot`` bit 0) writing the pair tensor the training
step consumes directly -- no fp32 NCHW tensors, no H2D of 5 x D x D floats per sample.
"""
import collections

import numpy as np
import torch

from . import _lib
from .engine import DATA_MEAN, DATA_STD

PairSpec = collections.namedtuple("PairSpec", "idx1 idx2 x y s flip swapped labels")


def _crop_box(boxes, idx1, idx2, base_aug, phase, randshift, rng):
    """_get_pair geometry (depth_occ_order_dataset.py:143-158), float64 with python int() truncation."""
    b = np.asarray(boxes, dtype=np.float64)[[idx1, idx2], :]
    l, u = b[:, 0].min(), b[:, 1].min()
    r, d = (b[:, 0] + b[:, 2]).max(), (b[:, 1] + b[:, 3]).max()
    w, h = r - l, d - u
    cx, cy = l + w / 2., u + h / 2.
    size = max([np.sqrt(w * h * 2.), w * 1.1, h * 1.1])
    if phase == "train":
        if randshift:
            cx += rng.uniform(*base_aug["shift"]) * size
            cy += rng.uniform(*base_aug["shift"]) * size
        size /= rng.uniform(*base_aug["scale"])
    return int(cx - size / 2.), int(cy - size / 2.), int(size)


def sample_pair(algo, boxes, gt, base_aug, pair=None, phase="train", extend_bidirec=False, rng=np.random,
                mode="patch"):
    """One training sample's random draws + labels, in the reference's order.  ``gt``: dict with the image's
    ``occ`` / ``depth`` / ``overlap`` / ``count`` matrices (reader.get_gt_ordering).  ``pair``: (idx1, idx2) for the
    depth datasets (they enumerate annotated pairs); None for the occlusion datasets, which draw it (70 % occluding
    pair / 30 % non-pair, occ_order_dataset.py:217-231).  Returns a PairSpec; labels are in ``set_input`` order."""
    label = None
    if pair is None:
        occ = np.array(gt["occ"]).copy()
        np.fill_diagonal(occ, -1)
        pairs, non_pairs = np.where(occ == 1), np.where(occ == 0)
        if len(pairs[0]) == 0:
            raise ValueError("image without an occluding pair (the reference re-draws another image)")
        if rng.rand() < 0.7 or len(non_pairs[0]) == 0:
            r = rng.choice(len(pairs[0]))
            idx1, idx2 = int(pairs[0][r]), int(pairs[1][r])
            label = 3 if (extend_bidirec and occ[idx2, idx1]) else 1
        else:
            r = rng.choice(len(non_pairs[0]))
            idx1, idx2 = int(non_pairs[0][r]), int(non_pairs[1][r])
            label = 2
    else:
        idx1, idx2 = int(pair[0]), int(pair[1])
    if mode == "patch":
        x, y, s = _crop_box(boxes, idx1, idx2, base_aug, phase, True, rng)
    else:                        # `resize` / `image` modes use the whole image: no crop jitter is drawn (:81-140)
        x, y, s = 0, 0, 0
    flip = bool(base_aug["flip"] and rng.rand() > 0.5)
    swapped = not (rng.rand() < 0.5)
    if algo in ("InstaOrderNet_od", "InstaOrderNet_d"):
        dm = gt["depth"]
        if dm[idx1, idx2] == -1:
            lab = -1
        elif dm[idx1, idx2] == 1 and dm[idx2, idx1] == 0:
            lab = 0
        elif dm[idx1, idx2] == 2:
            lab = 2
        else:
            raise ValueError("inconsistent depth annotation for (%d, %d)" % (idx1, idx2))
        if swapped and lab == 0:
            lab = 1
        labels = [lab, int(gt["count"][idx1, idx2]), int(gt["overlap"][idx1, idx2])]
        if algo == "InstaOrderNet_od":
            a_over_b, b_over_a = gt["occ"][idx1, idx2], gt["occ"][idx2, idx1]
            labels += [a_over_b, b_over_a] if swapped else [b_over_a, a_over_b]
    elif algo == "OrderNet":
        labels = [(0 if label == 1 else label) if swapped else label]
    else:   # InstaOrderNet_o
        a_over_b, b_over_a = gt["occ"][idx1, idx2], gt["occ"][idx2, idx1]
        labels = [a_over_b, b_over_a] if swapped else [b_over_a, a_over_b]
    return PairSpec(idx1, idx2, x, y, s, flip, swapped, labels)


class TrainBatchBuilder(object):
    """Builds the pair tensor + label tensors of a training batch on the device.  ``scenes``: objects with ``image``
    [H,W,3] u8, ``masks`` [N,H,W] u8 (instaorder_b200.engine.Scene); images and masks are uploaded once per scene and
    cached, so a batch costs one small descriptor upload + one kernel."""

    def __init__(self, algo, input_size, batch_pairs, device="cuda:0", mode="patch", max_scenes=64):
        self.lib = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("TrainBatchBuilder needs a CUDA device (there is no CPU fallback)")
        if mode not in ("patch", "resize", "image"):
            raise ValueError("patch_or_image %r" % (mode,))
        self.mode = mode
        self.algo, self.D, self.B = algo, int(input_size), int(batch_pairs)
        self.device = torch.device(device)
        self.pair_tensor = torch.zeros(int(self.lib.io_pair_tensor_bytes(self.B, self.D)), dtype=torch.uint8,
                                       device=self.device)
        self.mean = np.asarray(DATA_MEAN, dtype=np.float32)
        self.std = np.asarray(DATA_STD, dtype=np.float32)
        self._cache = {}
        self.gpu_launches = 0
        # whole-image modes: one pre-resized fp32 rgb plane per resident scene, addressed by slot
        self._planes = torch.empty((int(max_scenes), self.D, self.D, 3), dtype=torch.float32, device=self.device) \
            if mode != "patch" else None
        self._lut = torch.empty(768, dtype=torch.float32, device=self.device)

    def _resident(self, scene):
        key = id(scene)
        if key not in self._cache:
            img = torch.from_numpy(scene.image).to(self.device)
            msk = scene.masks_dev if getattr(scene, "masks_dev", None) is not None else \
                torch.from_numpy(scene.masks).to(self.device)
            slot = None
            if self.mode != "patch":     # the whole-image rgb is the same for every pair of the scene: resize it once
                slot = len(self._cache)
                if slot >= self._planes.shape[0]:
                    raise RuntimeError("TrainBatchBuilder: more than max_scenes = %d resident scenes" %
                                       self._planes.shape[0])
                h, w = scene.image.shape[:2]
                fn = self.lib.io_image_resize_linear_rgb if self.mode == "resize" else \
                    self.lib.io_image_square_linear_rgb
                _lib.check(fn(img.data_ptr(), h, w, self.D, _lib.ptr(self.mean), _lib.ptr(self.std),
                              self._lut.data_ptr(), self._planes[slot].data_ptr(), _lib.stream_ptr()))
                self.gpu_launches += 1
            self._cache[key] = (scene, img, msk, slot)
        return self._cache[key][1:]

    def build(self, scenes, specs):
        """scenes[k], specs[k] (PairSpec) for k < batch.  Returns (pair_tensor, labels [B, n_labels] float64 numpy)."""
        assert len(scenes) == len(specs) == self.B
        # one packed image / mask address space per call: offsets relative to the lowest base pointer
        res = [self._resident(s) for s in scenes]
        img_base = min(t[0].data_ptr() for t in res)
        msk_base = min(t[1].data_ptr() for t in res)
        desc = np.zeros(self.B, dtype=_lib.PAIR_DESC_DTYPE)
        if self.mode == "patch":
            for k, (sc, sp, (img, msk, _)) in enumerate(zip(scenes, specs, res)):
                if sp.s <= 0:
                    raise _lib.IoError(_lib.IO_ERR_DEGENERATE, "degenerate training pair (crop side %d)" % sp.s)
                n, h, w = (sc.n, sc.h, sc.w) if hasattr(sc, "n") else sc.masks.shape
                a, b = (sp.idx2, sp.idx1) if sp.swapped else (sp.idx1, sp.idx2)
                desc[k] = (img.data_ptr() - img_base, msk.data_ptr() - msk_base + a * h * w,
                           msk.data_ptr() - msk_base + b * h * w, h, w, sp.x, sp.y, sp.s, 1 if sp.flip else 0)
            d_desc = torch.from_numpy(desc.view(np.uint8)).to(self.device)
            _lib.check(self.lib.io_pair_gather_patch(img_base, msk_base, d_desc.data_ptr(), self.B, self.D,
                                                     _lib.ptr(self.mean), _lib.ptr(self.std),
                                                     self.pair_tensor.data_ptr(), _lib.stream_ptr()))
        else:
            for k, (sc, sp, (img, msk, slot)) in enumerate(zip(scenes, specs, res)):
                n, h, w = (sc.n, sc.h, sc.w) if hasattr(sc, "n") else sc.masks.shape
                a, b = (sp.idx2, sp.idx1) if sp.swapped else (sp.idx1, sp.idx2)
                sq = max(h, w) if self.mode == "image" else 0
                desc[k] = (0, msk.data_ptr() - msk_base + a * h * w, msk.data_ptr() - msk_base + b * h * w, h, w,
                           (sq - w) // 2 if sq else 0, (sq - h) // 2 if sq else 0, sq,
                           slot | (0x40000000 if sp.flip else 0))
            d_desc = torch.from_numpy(desc.view(np.uint8)).to(self.device)
            _lib.check(self.lib.io_pair_gather_resize(self._planes.data_ptr(), msk_base, d_desc.data_ptr(), self.B,
                                                      self.D, self.pair_tensor.data_ptr(), _lib.stream_ptr()))
        self.gpu_launches += 1
        labels = np.asarray([sp.labels for sp in specs], dtype=np.float64)
        return self.pair_tensor, labels
