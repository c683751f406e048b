"""Host side of the B200 pairwise-order path: batches the instance pairs of many images, stages the u8 images /
masks through pinned memory, and drives the CUDA kernels behind the C ABI (``include/instaorder_b200.h``).

The reference runs ``inference.py:349-624`` one pair at a time (2 batch-1 forwards + 5 host syncs per pair).  Here a
*batch* is up to ``max_pairs`` pairs taken from consecutive images; per batch the device runs

    H2D(images, masks, descriptors) -> fused gather -> ResNet-50 (both directions) -> decide + scatter

with no host synchronisation; the order matrices of the whole call come back in one D2H copy at the end.
PyTorch is used for device memory, pinned staging buffers and streams only.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

DATA_MEAN = (0.485, 0.456, 0.406)   # reference utils/data_utils.py:9-10
DATA_STD = (0.229, 0.224, 0.225)

# algo -> list of (head kind, number of logits, writes 'occ' | 'depth')
HEADS = {
    "InstaOrderNet_o": [(_lib.IO_HEAD_OCC, 2, "occ")],
    "InstaOrderNet_d": [(_lib.IO_HEAD_DEPTH, 3, "depth")],
    "InstaOrderNet_od": [(_lib.IO_HEAD_OCC, 2, "occ"), (_lib.IO_HEAD_DEPTH, 3, "depth")],
    "OrderNet": [(_lib.IO_HEAD_ORDERNET, None, "occ")],   # 3 or 4 logits (OrderNet_ext)
}


def heads_for(algo, num_classes):
    if algo not in HEADS:
        raise ValueError("method name should be one of %s" % sorted(HEADS))
    ncs = list(num_classes) if isinstance(num_classes, (list, tuple)) else [int(num_classes)]
    heads = []
    for (kind, k, what), nc in zip(HEADS[algo], ncs):
        k = nc if k is None else k
        if k != nc:
            raise ValueError("%s expects %d logits, checkpoint head has %d" % (algo, k, nc))
        heads.append((kind, k, what))
    if len(heads) != len(ncs):
        raise ValueError("%s: num_classes %r does not match its heads" % (algo, num_classes))
    return heads


class Scene:
    """One image with its instances: the input contract of ``infer_order_sup_*`` (image [H,W,3] u8, inmodal
    [N,H,W] u8, bboxes [N,4] xywh)."""
    __slots__ = ("image", "masks", "masks_dev", "boxes", "n", "h", "w")

    def __init__(self, image, masks, boxes):
        self.image = np.ascontiguousarray(image, dtype=np.uint8)
        if isinstance(masks, torch.Tensor) and masks.is_cuda:
            # masks already in HBM (instaorder_b200.masks.rasterize): staged device-to-device, never through the host
            if masks.dtype != torch.uint8 or masks.dim() != 3:
                raise ValueError("device masks must be uint8 [N, H, W]")
            self.masks, self.masks_dev = None, masks.contiguous()
            self.n, self.h, self.w = (int(v) for v in masks.shape)
        else:
            self.masks, self.masks_dev = np.ascontiguousarray(masks, dtype=np.uint8), None
            self.n, self.h, self.w = self.masks.shape
        self.boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
        if self.image.shape != (self.h, self.w, 3):
            raise ValueError("image %r does not match masks %r" % (self.image.shape, (self.n, self.h, self.w)))
        if self.boxes.shape[0] != self.n:
            raise ValueError("%d boxes for %d masks" % (self.boxes.shape[0], self.n))


def enumerate_pairs(n):
    out = np.empty((n * (n - 1) // 2, 2), dtype=np.int32)
    c = _lib.check(_lib.lib().io_pair_enumerate(n, _lib.ptr(out)))
    assert c == out.shape[0]
    return out


def expand_bbox(bboxes, enlarge_box=3.0):
    """``Tester.expand_bbox`` (reference tools/test.py:155-163)."""
    b = np.ascontiguousarray(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4))
    out = np.empty((b.shape[0], 4), dtype=np.int32)
    _lib.check(_lib.lib().io_expand_bbox(_lib.ptr(b), b.shape[0], float(enlarge_box), _lib.ptr(out)))
    return out.astype(np.int64)


def closest_multiple_of(n, m=32):
    """``get_closest_int_multiple_of`` (reference utils/data_utils.py:13-17): ties go up."""
    r = n % m
    return n + m - r if r >= m // 2 else n - r


def pair_crop_boxes(boxes, pairs):
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    out = np.empty((pairs.shape[0], 4), dtype=np.int32)
    _lib.check(_lib.lib().io_pair_crop_boxes(_lib.ptr(b), _lib.ptr(pairs), pairs.shape[0], _lib.ptr(out)))
    return out


class _Slot:
    """Pinned staging buffers + their device twins for one in-flight batch."""

    def __init__(self, img_bytes, mask_bytes, max_pairs, device):
        self.h_img = torch.empty(img_bytes, dtype=torch.uint8).pin_memory()
        self.h_mask = torch.empty(mask_bytes, dtype=torch.uint8).pin_memory()
        self.h_desc = torch.empty(max_pairs * 48, dtype=torch.uint8).pin_memory()
        # rows of max_pairs int64: [0] = (i, j) as 2 x int32, [1] = matrix side N as int32, [2] = matrix offset
        self.h_meta = torch.empty(max_pairs * 4, dtype=torch.int64).pin_memory()
        self.d_img = torch.empty(img_bytes, dtype=torch.uint8, device=device)
        self.d_mask = torch.empty(mask_bytes, dtype=torch.uint8, device=device)
        self.d_desc = torch.empty(max_pairs * 48, dtype=torch.uint8, device=device)
        self.d_meta = torch.empty(max_pairs * 4, dtype=torch.int64, device=device)
        self.event = None        # compute that consumed the device buffers has finished
        self.copied = None       # H2D copies out of the pinned buffers have finished


class _Pending:
    """Handle of a submitted call (OrderEngine.submit_scenes): pinned result buffers + the event behind their D2H."""
    __slots__ = ("scenes", "mat_offs", "bufs", "event")

    def __init__(self, scenes, mat_offs, bufs, event):
        self.scenes, self.mat_offs, self.bufs, self.event = scenes, mat_offs, bufs, event


class _Resident:
    """Device-resident inputs of one batch (same attribute names as _Slot where the kernels need them)."""
    pass


_PACK_POOL = None
_PACK_THREADS = int(os.environ.get("INSTAORDER_PACK_THREADS", max(1, min(4, (os.cpu_count() or 4) // 4))))


def _copy_one(job):
    dst, off, nbytes, src = job
    dst[off:off + nbytes] = src.reshape(-1)


def _pack_copies(copies):
    """Byte copies of a batch's images / masks into the pinned staging slot, spread over a small thread pool."""
    global _PACK_POOL
    if _PACK_THREADS <= 1 or len(copies) < 2:
        for job in copies:
            _copy_one(job)
        return
    if _PACK_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _PACK_POOL = ThreadPoolExecutor(max_workers=_PACK_THREADS, thread_name_prefix="io-pack")
    list(_PACK_POOL.map(_copy_one, copies))


class OrderEngine:
    def __init__(self, num_classes, input_size=256, max_pairs=256, device="cuda:0", data_mean=DATA_MEAN,
                 data_std=DATA_STD, img_bytes=32 << 20, mask_bytes=256 << 20, slots=2):
        if not torch.cuda.is_available():
            raise RuntimeError("instaorder_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.ncs = list(num_classes) if isinstance(num_classes, (list, tuple)) else [int(num_classes)]
        self.k_total = sum(self.ncs)
        self.d = int(input_size)
        self.max_pairs = int(max_pairs)
        self.first_batch_pairs = max(1, min(self.max_pairs, int(os.environ.get("INSTAORDER_FIRST_BATCH", "64"))))
        self.mean = np.asarray(data_mean, dtype=np.float32)
        self.std = np.asarray(data_std, dtype=np.float32)
        self.net = None
        self.max_items_per_batch = 1 << 30    # images (or image parts) per batch; the InstaDepthNet engine caps it
        self._create_nets()
        self.pair_tensor = torch.zeros(self.lib.io_pair_tensor_bytes(self.max_pairs, self.d), dtype=torch.uint8,
                                       device=self.device)
        self.logits = torch.empty((self.max_pairs, 2, self.k_total), dtype=torch.float32, device=self.device)
        self.margins = torch.empty((2, self.max_pairs), dtype=torch.float32, device=self.device)
        self._slot_args = (img_bytes, mask_bytes, self.max_pairs, self.device)
        self._slots = [None] * slots
        self._slot_i = 0
        self.copy_stream = torch.cuda.Stream(device=self.device)   # H2D of batch k+1 overlaps compute of batch k
        self.gpu_launches = 0   # kernels launched by this engine since creation
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _create_nets(self):
        ncs = np.asarray(self.ncs, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(self.lib.io_net_create(_lib.ptr(ncs), len(self.ncs), self.d, self.max_pairs, C.byref(h)))
        self.net = h

    def __del__(self):
        try:
            if getattr(self, "net", None):
                self.lib.io_net_destroy(self.net)
                self.net = None
        except Exception:
            pass

    # ---- weights ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """``sd``: a reference ``state_dict`` (keys with or without ``module.``; torch tensors or numpy)."""
        names, arrs = [], []
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                continue
            k = k[7:] if k.startswith("module.") else k
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            names.append(k.encode())
            arrs.append(np.ascontiguousarray(a, dtype=np.float32))
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_numel = (C.c_int64 * n)(*[a.size for a in arrs])
        _lib.check(self.lib.io_net_load_state(self.net, c_names, c_ptrs, c_numel, n))

    # ---- low-level steps (also used by the tests) ------------------------------------------------------------
    def _ensure_capacity(self, img_bytes, mask_bytes):
        """Grows the pinned / device staging slots when ONE scene alone does not fit them (a very large image, or
        many instances at high resolution).  Slots in flight are drained first; happens at most a few times per
        engine, never in the steady state."""
        cur_i, cur_m = self._slot_args[0], self._slot_args[1]
        if img_bytes <= cur_i and mask_bytes <= cur_m:
            return
        torch.cuda.synchronize(self.device)
        grow = lambda cur, need: cur if need <= cur else max(need, 2 * cur)
        self._slot_args = (grow(cur_i, img_bytes), grow(cur_m, mask_bytes)) + tuple(self._slot_args[2:])
        self._slots = [None] * len(self._slots)
        self._slot_i = 0

    @staticmethod
    def _scene_bytes(sc):
        return (sc.h * sc.w * 3 + 15) // 16 * 16, (sc.n * sc.h * sc.w + 15) // 16 * 16

    def _slot(self):
        i = self._slot_i
        self._slot_i = (i + 1) % len(self._slots)
        if self._slots[i] is None:
            self._slots[i] = _Slot(*self._slot_args)
        s = self._slots[i]
        if s.copied is not None:
            s.copied.synchronize()  # the pinned buffers have been read by the previous H2D that used this slot
        return s

    def stage_batch(self, items, mode="patch"):
        """items: list of (scene, pairs int32[p,2], crops int32[p,4] | None, mat_off, scene_index).
        Packs images / masks / descriptors into a pinned slot and issues the H2D copies.  Returns (slot, P)."""
        s = self._slot()
        img_off = mask_off = 0
        P = 0
        desc = np.frombuffer(s.h_desc.numpy(), dtype=_lib.PAIR_DESC_DTYPE)
        meta = s.h_meta.numpy().reshape(4, self.max_pairs)
        ij = meta[0:1].view(np.int32).reshape(-1)[: 2 * self.max_pairs].reshape(self.max_pairs, 2)
        n32 = meta[1:2].view(np.int32).reshape(-1)[: self.max_pairs]
        himg, hmask = s.h_img.numpy(), s.h_mask.numpy()
        slot_idx = 0
        self.resize_jobs = []
        dev_masks = []
        copies = []
        for (sc, pairs, crops, mat_off, _) in items:
            p = pairs.shape[0]
            ib, mb = sc.h * sc.w * 3, sc.n * sc.h * sc.w
            if img_off + ib > himg.size or mask_off + mb > hmask.size:
                raise ValueError("staging buffers too small for this batch: it needs more than %d B of images / %d B "
                                 "of masks (this scene: %d / %d B).  infer_scenes() cuts batches to fit; callers of "
                                 "stage_batch() / make_batches() must pass fewer scenes per batch or create the "
                                 "engine with larger img_bytes= / mask_bytes=" % (himg.size, hmask.size, ib, mb))
            # the byte copies into the pinned slot (65 MB per 765-pair call of the bench workload) run on a few threads:
            # numpy releases the GIL inside them, and a single thread's ~10 GB/s was the un-overlapped part of a call
            copies.append((himg, img_off, ib, sc.image))
            if sc.masks_dev is None:
                copies.append((hmask, mask_off, mb, sc.masks))
            else:
                dev_masks.append((mask_off, mb, sc.masks_dev))
            d = desc[P:P + p]
            d["image_off"] = img_off
            d["mask_a_off"] = mask_off + pairs[:, 0].astype(np.int64) * (sc.h * sc.w)
            d["mask_b_off"] = mask_off + pairs[:, 1].astype(np.int64) * (sc.h * sc.w)
            d["h"], d["w"] = sc.h, sc.w
            if crops is not None:
                d["x"], d["y"], d["s"] = crops[:, 0], crops[:, 1], crops[:, 2]
            elif mode == "image":      # centred zero-padded square, reference inference.py:378-382
                sq = max(sc.h, sc.w)
                d["x"], d["y"], d["s"] = (sq - sc.w) // 2, (sq - sc.h) // 2, sq
            else:
                d["x"] = d["y"] = d["s"] = 0
            d["rgb_slot"] = slot_idx
            if mode in ("resize", "image", "orig"):
                self.resize_jobs.append((img_off, sc.h, sc.w, slot_idx))
                slot_idx += 1
            ij[P:P + p] = pairs
            meta[2, P:P + p] = mat_off
            n32[P:P + p] = sc.n
            img_off += (ib + 15) // 16 * 16
            mask_off += (mb + 15) // 16 * 16
            P += p
        _pack_copies(copies)
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if s.event is not None:
                self.copy_stream.wait_event(s.event)   # kernels of the batch that last used these device buffers
            s.d_img[:img_off].copy_(s.h_img[:img_off], non_blocking=True)
            if len(dev_masks) < len(items):
                s.d_mask[:mask_off].copy_(s.h_mask[:mask_off], non_blocking=True)
            if dev_masks:
                self.copy_stream.wait_stream(compute)     # the kernels that produced the device masks
                for (o, nb, t) in dev_masks:
                    s.d_mask[o:o + nb].copy_(t.reshape(-1), non_blocking=True)
            s.d_desc[:P * 48].copy_(s.h_desc[:P * 48], non_blocking=True)
            s.d_meta.copy_(s.h_meta, non_blocking=True)
            s.copied = torch.cuda.Event()
            s.copied.record(self.copy_stream)
        compute.wait_event(s.copied)
        self.h2d_bytes += img_off + mask_off - sum(nb for (_, nb, _) in dev_masks) + P * 48 + s.h_meta.numel() * 8
        return s, P

    def gather(self, s, P, mode="patch", geom=None):
        st = _lib.stream_ptr()
        if mode == "orig":
            # reference inference.py:401-408: the image at its own size rounded to multiples of 32 (geom = (hh, ww)),
            # transform_resize for the rgb, cv2.INTER_NEAREST for both masks; ONE image per batch
            hh, ww = geom
            (img_off, h, w, slot), = self.resize_jobs
            need = hh * ww * 3
            if getattr(self, "_planes", None) is None or self._planes.numel() < need:
                self._planes = torch.empty(need, dtype=torch.float32, device=self.device)
                self._lut_scratch = torch.empty(768, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.io_image_resize_rgb_hw(s.d_img.data_ptr() + img_off, h, w, hh, ww, _lib.ptr(self.mean),
                                                       _lib.ptr(self.std), self._planes.data_ptr(), st))
            _lib.check(self.lib.io_pair_gather_resize_hw(self._planes.data_ptr(), s.d_mask.data_ptr(),
                                                         s.d_desc.data_ptr(), P, hh, ww, self.pair_tensor.data_ptr(), st))
            self.gpu_launches += 2
            return
        if mode == "patch":
            _lib.check(self.lib.io_pair_gather_patch(s.d_img.data_ptr(), s.d_mask.data_ptr(), s.d_desc.data_ptr(), P,
                                                     self.d, _lib.ptr(self.mean), _lib.ptr(self.std),
                                                     self.pair_tensor.data_ptr(), st))
            self.gpu_launches += 1
        elif mode in ("resize", "image"):
            n_img = len(self.resize_jobs)
            need = n_img * self.d * self.d * 3
            if getattr(self, "_planes", None) is None or self._planes.numel() < need:
                self._planes = torch.empty(need, dtype=torch.float32, device=self.device)
                self._lut_scratch = torch.empty(768, dtype=torch.float32, device=self.device)
            for (img_off, h, w, slot) in self.resize_jobs:
                dst = self._planes.data_ptr() + slot * self.d * self.d * 12
                if mode == "resize":
                    _lib.check(self.lib.io_image_resize_rgb(s.d_img.data_ptr() + img_off, h, w, self.d,
                                                            _lib.ptr(self.mean), _lib.ptr(self.std), dst, st))
                else:
                    _lib.check(self.lib.io_image_square_linear_rgb(s.d_img.data_ptr() + img_off, h, w, self.d,
                                                                   _lib.ptr(self.mean), _lib.ptr(self.std),
                                                                   self._lut_scratch.data_ptr(), dst, st))
            _lib.check(self.lib.io_pair_gather_resize(self._planes.data_ptr(), s.d_mask.data_ptr(),
                                                      s.d_desc.data_ptr(), P, self.d, self.pair_tensor.data_ptr(),
                                                      st))
            self.gpu_launches += n_img + 1
        else:
            raise NotImplementedError("patch_or_image=%r (supported: 'patch', 'resize', 'image')" % (mode,))

    def forward(self, P, geom=None):
        if geom is not None:      # `orig` mode: non-square network input of this image
            _lib.check(self.lib.io_net_forward_pairs_hw(self.net, self.pair_tensor.data_ptr(), P, geom[0], geom[1],
                                                        self.logits.data_ptr(), _lib.stream_ptr()))
        else:
            _lib.check(self.lib.io_net_forward_pairs(self.net, self.pair_tensor.data_ptr(), P, self.logits.data_ptr(),
                                                     _lib.stream_ptr()))
        self.gpu_launches += self.lib.io_net_last_launches(self.net)

    def decide(self, s, P, heads, mats):
        """heads: [(kind, k, 'occ'|'depth')]; mats: dict name -> flat int64 device tensor."""
        meta = s.d_meta.view(4, self.max_pairs)
        off = 0
        for hi, (kind, k, what) in enumerate(heads):
            _lib.check(self.lib.io_order_decide(self.logits.data_ptr(), P, self.k_total, kind, off, k,
                                                meta[0].data_ptr(), meta[2].data_ptr(), meta[1].data_ptr(),
                                                mats[what].data_ptr(), self.margins[hi].data_ptr(),
                                                _lib.stream_ptr()))
            off += k
            self.gpu_launches += 1

    def finish(self, s):
        s.event = torch.cuda.Event()
        s.event.record()

    # ---- the batched driver ------------------------------------------------------------------------------
    def infer_scenes(self, scenes, algo, pairs="all", patch_or_image="patch", return_details=False, _pending=False,
                     _first_cap=None):
        """Order matrices for a list of ``Scene``.  Returns a list of dicts with 'occ' / 'depth' int64 [N,N]
        (+ 'pairs', 'logits', 'margin_occ', 'margin_depth' when ``return_details``)."""
        heads = heads_for(algo, self.ncs if len(self.ncs) > 1 else self.ncs[0])
        mode = patch_or_image
        if mode not in ("patch", "resize", "image", "orig"):
            raise ValueError("patch_or_image=%r (one of 'patch', 'resize', 'image', 'orig')" % (mode,))
        if pairs not in ("all", "nbor"):
            raise ValueError("pairs must be 'all' or 'nbor'")
        mat_offs = []
        tot = total_pairs = 0
        for sc in scenes:
            mat_offs.append(tot)
            tot += sc.n * sc.n
            total_pairs += sc.n * (sc.n - 1) // 2
        mats = {w: torch.zeros(max(tot, 1), dtype=torch.int64, device=self.device) for (_, _, w) in heads}
        details = [dict(pairs=None, logits=[], margins=[]) for _ in scenes] if return_details else None
        # batches of <= max_pairs pairs; pairs + crop windows (host, float64 -- bit-exact with the reference's
        # geometry) are computed scene by scene while earlier batches already run on the GPU
        batch, count = [], 0
        used = [0, 0]          # image / mask bytes the current batch needs in its staging slot

        def flush():
            nonlocal batch, count
            if not batch:
                return
            used[0] = used[1] = 0
            geom = None
            if mode == "orig":     # one image per batch: its own size rounded to multiples of 32 is the network input
                sc0 = batch[0][0]
                geom = (closest_multiple_of(sc0.h), closest_multiple_of(sc0.w))
                if min(geom) < 32 or max(geom) > self.d:
                    raise ValueError("'orig' mode: image %d x %d -> network input %d x %d, this engine handles 32 .. %d "
                                     "(models.engine_for_orig sizes the engine from the image)" %
                                     (sc0.h, sc0.w, geom[0], geom[1], self.d))
            s, P = self.stage_batch(batch, mode)
            self.gather(s, P, mode, geom)
            self.forward(P, geom)
            self.decide(s, P, heads, mats)
            self.finish(s)
            if return_details:
                lg = self.logits[:P].cpu().numpy()
                mg = self.margins[:, :P].cpu().numpy()
                self.d2h_bytes += lg.nbytes + mg.nbytes
                o = 0
                for (_, pr, _, _, si) in batch:
                    details[si]["logits"].append(lg[o:o + pr.shape[0]])
                    details[si]["margins"].append(mg[:, o:o + pr.shape[0]])
                    o += pr.shape[0]
            batch, count = [], 0

        # the first batch of a long call is kept short: the GPU starts while the host is still packing the second one
        # (staging = a host memcpy into pinned memory, ~0.1 ms per MB, otherwise fully exposed at the start of the call)
        cap = self.first_batch_pairs if total_pairs > self.max_pairs else self.max_pairs
        if _first_cap is not None:
            cap = _first_cap
        for si, sc in enumerate(scenes):
            pr = enumerate_pairs(sc.n)
            if pairs == "nbor" and pr.shape[0]:
                pr = pr[self.bordering(sc, pr)]
            crops = pair_crop_boxes(sc.boxes, pr) if (mode == "patch" and pr.shape[0]) else None
            if return_details:
                details[si]["pairs"] = pr
            o = 0
            while o < pr.shape[0]:
                take = min(pr.shape[0] - o, cap - count)
                ib, mb = self._scene_bytes(sc)
                fits = used[0] + ib <= self._slot_args[0] and used[1] + mb <= self._slot_args[1]
                if take == 0 or len(batch) >= (1 if mode == "orig" else self.max_items_per_batch) or (batch and not fits):
                    flush()                      # pair budget, item budget or staging bytes exhausted: start a new batch
                    cap = self.max_pairs
                    continue
                if not fits:                     # a single scene larger than a slot: grow the slots once
                    self._ensure_capacity(ib, mb)
                batch.append((sc, pr[o:o + take], None if crops is None else crops[o:o + take], mat_offs[si], si))
                used[0] += ib
                used[1] += mb
                count += take
                o += take
                if count == cap:
                    flush()
                    cap = self.max_pairs
        flush()
        # 3. one D2H for every matrix of the call
        if _pending:       # submit_scenes(): the copy goes to pinned memory behind the kernels; collect() waits for it
            bufs = {}
            for w, m in mats.items():
                hb = self._pinned_i64(m.numel())
                hb.copy_(m, non_blocking=True)
                bufs[w] = hb
            ev = torch.cuda.Event()
            ev.record()
            return _Pending(scenes, mat_offs, bufs, ev)
        host = {w: m.cpu().numpy() for w, m in mats.items()}
        self.d2h_bytes += sum(v.nbytes for v in host.values())
        out = []
        for si, sc in enumerate(scenes):
            r = {}
            for w, m in host.items():
                r[w] = m[mat_offs[si]:mat_offs[si] + sc.n * sc.n].reshape(sc.n, sc.n).copy()
            if return_details:
                d = details[si]
                r["pairs"] = d["pairs"]
                r["logits"] = np.concatenate(d["logits"]) if d["logits"] else np.zeros((0, 2, self.k_total), np.float32)
                mg = np.concatenate(d["margins"], axis=1) if d["margins"] else np.zeros((2, 0), np.float32)
                for hi, (_, _, w) in enumerate(heads):
                    r["margin_" + w] = mg[hi]
            out.append(r)
        return out

    # ---- pipelined calls: the next call's packing / H2D runs while this call's kernels execute ----------------
    def _pinned_i64(self, n):
        pool = self.__dict__.setdefault("_pin_pool", {})
        free = pool.setdefault(n, [])
        return free.pop() if free else torch.empty(n, dtype=torch.int64).pin_memory()

    def submit_scenes(self, scenes, algo, pairs="all", patch_or_image="patch", first_batch_pairs=None):
        """``infer_scenes`` without the final wait: everything is enqueued (incl. the D2H of the order matrices into
        pinned memory) and a handle comes back; ``collect(handle)`` returns what ``infer_scenes`` would have.  Several
        calls may be in flight (the pinned staging slots are recycled behind CUDA events)."""
        return self.infer_scenes(scenes, algo, pairs, patch_or_image, False, _pending=True, _first_cap=first_batch_pairs)

    def collect(self, pending):
        pending.event.synchronize()
        out = []
        nbytes = 0
        for si, sc in enumerate(pending.scenes):
            r = {}
            for w, hb in pending.bufs.items():
                o = pending.mat_offs[si]
                r[w] = hb[o:o + sc.n * sc.n].numpy().reshape(sc.n, sc.n).copy()
            out.append(r)
        for hb in pending.bufs.values():
            nbytes += hb.numel() * 8
            self._pin_pool[hb.numel()].append(hb)
        self.d2h_bytes += nbytes
        return out

    def infer_stream(self, calls, algo, pairs="all", patch_or_image="patch", depth=2):
        """Generator over an iterable of scene lists (one list = one ``infer_scenes`` call): yields each call's result in
        order while up to ``depth`` calls are in flight, so the host-side packing and the H2D copies of call k + 1 overlap
        the kernels of call k.  Same results as calling ``infer_scenes`` on every list."""
        queue = []
        for scenes in calls:
            # with work already in flight there is no idle GPU to feed early: full-size first batch
            queue.append(self.submit_scenes(scenes, algo, pairs, patch_or_image,
                                            first_batch_pairs=self.max_pairs if queue else None))
            if len(queue) >= depth:
                yield self.collect(queue.pop(0))
        while queue:
            yield self.collect(queue.pop(0))

    # ---- device-resident batches (bench.py `value`: inputs already in HBM when the timed region starts) -------
    def make_batches(self, scenes, pairs_per_batch=None, mode="patch"):
        """Cuts the pair stream of ``scenes`` into batches of exactly ``pairs_per_batch`` pairs (the last, short
        batch is dropped).  Returns a list of item lists for ``stage_batch``."""
        ppb = pairs_per_batch or self.max_pairs
        batches, cur, count, mat_off = [], [], 0, 0
        for si, sc in enumerate(scenes):
            pr = enumerate_pairs(sc.n)
            crops = pair_crop_boxes(sc.boxes, pr) if mode == "patch" else None
            o = 0
            while o < pr.shape[0]:
                take = min(pr.shape[0] - o, ppb - count)
                cur.append((sc, pr[o:o + take], None if crops is None else crops[o:o + take], mat_off, si))
                count += take
                o += take
                if count == ppb:
                    batches.append(cur)
                    cur, count = [], 0
            mat_off += sc.n * sc.n
        return batches, mat_off

    def upload_resident(self, items, mat_elems, mode="patch"):
        """Stages one batch and keeps private device copies of its inputs."""
        s, P = self.stage_batch(items, mode)
        torch.cuda.current_stream().synchronize()
        r = _Resident()
        r.P = P
        img_bytes = sum((it[0].h * it[0].w * 3 + 15) // 16 * 16 for it in items)
        mask_bytes = sum((it[0].n * it[0].h * it[0].w + 15) // 16 * 16 for it in items)
        r.d_img = s.d_img[:img_bytes].clone()
        r.d_mask = s.d_mask[:mask_bytes].clone()
        r.d_desc = s.d_desc.clone()
        r.d_meta = s.d_meta.clone()
        r.resize_jobs = list(self.resize_jobs)
        r.input_bytes = img_bytes + mask_bytes
        r.mats = None
        r.mat_elems = mat_elems
        return r

    def run_resident(self, r, heads, mode="patch"):
        """gather -> forward -> decide on a resident batch; no host synchronisation, no copies."""
        if r.mats is None:
            r.mats = {w: torch.zeros(max(r.mat_elems, 1), dtype=torch.int64, device=self.device) for (_, _, w) in heads}
        self.resize_jobs = r.resize_jobs
        self.gather(r, r.P, mode)
        self.forward(r.P)
        self.decide(r, r.P, heads, r.mats)

    def bordering(self, sc, pairs):
        """``bordering`` (reference inference.py:691-696) for every candidate pair of one scene -> bool[p]."""
        m = sc.masks_dev if sc.masks_dev is not None else torch.from_numpy(sc.masks).to(self.device)
        pr = torch.from_numpy(np.ascontiguousarray(pairs, dtype=np.int32)).to(self.device)
        flags = torch.empty(pairs.shape[0], dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.io_pair_bordering(m.data_ptr(), sc.n, sc.h, sc.w, pr.data_ptr(), pairs.shape[0],
                                              flags.data_ptr(), _lib.stream_ptr()))
        self.gpu_launches += 1
        self.h2d_bytes += m.numel() + pr.numel() * 4
        self.d2h_bytes += flags.numel()
        return flags.cpu().numpy().astype(bool)


# ---- metrics -----------------------------------------------------------------------------------------------
def _pack_mats(mats, device):
    ns = np.array([m.shape[0] for m in mats], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)[:-1]]).astype(np.int64)
    flat = np.concatenate([np.asarray(m, dtype=np.int64).reshape(-1) for m in mats]) if len(mats) else \
        np.zeros(0, np.int64)
    return torch.from_numpy(flat).to(device), torch.from_numpy(offs).to(device), torch.from_numpy(ns).to(device)


def _metric_device(device):
    """The GPU the metric kernels run on: the caller's, else the process's current device (``cuda:LOCAL_RANK`` once
    an engine exists) -- never a fixed ``cuda:0`` that another rank's stream would be launched against."""
    if not torch.cuda.is_available():
        raise RuntimeError("instaorder_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def metrics_prf(orders, gts, zd, device=None):
    """Batched ``eval_order_recall_precision_f1`` (reference inference.py:794-802) -> float64 [B, 3]."""
    device = _metric_device(device)
    with torch.cuda.device(device):          # stream and pointers of the same GPU
        o, off, ns = _pack_mats(orders, device)
        g, _, _ = _pack_mats(gts, device)
        out = torch.empty((len(orders), 3), dtype=torch.float64, device=device)
        _lib.check(_lib.lib().io_metrics_prf(o.data_ptr(), g.data_ptr(), off.data_ptr(), ns.data_ptr(), len(orders),
                                             int(zd), out.data_ptr(), _lib.stream_ptr()))
        return out.cpu().numpy()


def metrics_whdr(orders, gt_orders, gt_overlaps, gt_counts, device=None):
    """Batched ``eval_depth_order_whdr`` (reference inference.py:764-791) -> float64 [B, 9]."""
    device = _metric_device(device)
    with torch.cuda.device(device):
        o, off, ns = _pack_mats(orders, device)
        g, _, _ = _pack_mats(gt_orders, device)
        v, _, _ = _pack_mats(gt_overlaps, device)
        c, _, _ = _pack_mats(gt_counts, device)
        out = torch.empty((len(orders), 9), dtype=torch.float64, device=device)
        _lib.check(_lib.lib().io_metrics_whdr(o.data_ptr(), g.data_ptr(), v.data_ptr(), c.data_ptr(), off.data_ptr(),
                                              ns.data_ptr(), len(orders), out.data_ptr(), _lib.stream_ptr()))
        return out.cpu().numpy()
