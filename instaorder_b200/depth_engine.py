"""InstaDepthNet^od order inference (BASELINE config 5; reference ``midas/midas_net.py:113-212`` through
``inference.infer_order_sup_occ_depth(method="InstaDepthNet_od")``, ``inference.py:349-436, 107-137``).

The order outputs depend on the ResNeXt-101 32x8d encoder's layer1..3 -- a function of the IMAGE only -- and on the two
2-channel ResNet-50 trunks ``do_net`` / ``oo_net`` per pair direction, with the encoder features added in front of
their layer2 / layer3 / layer4.  The reference recomputes the encoder (and the whole MiDaS decoder, which only feeds the
disparity output) for every pair direction; here the encoder runs ONCE PER IMAGE of a batch and its three feature
maps are broadcast into the trunks by index (``io_net_set_inject``).  Everything runs through the kernels of the
pairwise-order path: three ``io_net_t`` handles created with ``io_net_create_arch`` -- the encoder's grouped 3x3
convolutions as dense block-diagonal weights, the 3- / 2-channel stems embedded in the 5-channel pair-tensor stem.
The disparity branch (encoder layer4 + ``scratch.*`` = the MiDaS decoder, ``midas_net.py:189-198``) is evaluated on
request (``with_disparity=True``: ``disparity()``, ``disparity_order()``), per image, through ``io_conv_bn_act`` +
``io_add_relu`` + ``io_upsample2x_bilinear``."""
import ctypes as C

import numpy as np
import torch

from . import _lib, synth
from .engine import OrderEngine, enumerate_pairs

RESNET50 = ((64, 128, 256, 512), (256, 512, 1024, 2048), (3, 4, 6, 3))


def _i32(v):
    return np.asarray(v, dtype=np.int32)


class DepthOrderEngine(OrderEngine):
    def __init__(self, input_size=384, max_pairs=64, max_images=16, device="cuda:0", with_occ=True,
                 with_disparity=False, with_trunks=True, **kw):
        """``with_occ=False``: InstaDepthNet^d (midas_net.py:15-110) -- no ``oo_net``, depth order only.
        ``with_disparity=True``: the encoder also runs layer4 and ``disparity()`` evaluates the MiDaS decoder."""
        self.max_images = int(max_images)
        self.with_occ = bool(with_occ)
        self.with_disparity = bool(with_disparity)
        self.with_trunks = bool(with_trunks)     # False: a plain MiDaS network (encoder + decoder), disparity only
        self._dec = None
        super().__init__([2, 3], input_size, max_pairs, device, **kw)

    # ---- handles -------------------------------------------------------------------------------------------------
    def _create_arch(self, arch, n_layers, keep, ncs, pairs):
        h = C.c_void_p()
        w, o, b = (_i32(a) for a in arch)
        nc = _i32(ncs) if ncs else None
        _lib.check(self.lib.io_net_create_arch(_lib.ptr(w), _lib.ptr(o), _lib.ptr(b), n_layers, int(keep),
                                               _lib.ptr(nc) if nc is not None else None, len(ncs) if ncs else 0,
                                               self.d, pairs, C.byref(h)))
        return h

    def _create_nets(self):
        self.max_items_per_batch = self.max_images      # the encoder handle is sized for this many images per batch
        self.enc_layers = 4 if self.with_disparity else 3
        self.enc = self._create_arch((synth.RESNEXT_WIDTHS, synth.RESNEXT_OUTS, synth.RESNEXT_BLOCKS), self.enc_layers,
                                     True, None, self.max_images)
        self.do_net = self._create_arch(RESNET50, 4, False, [3], self.max_pairs) if self.with_trunks else None
        self.oo_net = self._create_arch(RESNET50, 4, False, [2], self.max_pairs) \
            if (self.with_occ and self.with_trunks) else None
        self.inject_idx = torch.zeros(2 * self.max_pairs, dtype=torch.int32, device=self.device)
        feats = []
        for li in range(self.enc_layers):
            ptr, n = C.c_void_p(), C.c_int64()
            _lib.check(self.lib.io_net_feature(self.enc, li, C.byref(ptr), C.byref(n)))
            feats.append(ptr)
        self._enc_feats = feats
        for net in (self.do_net, self.oo_net):
            if net is None:
                continue
            _lib.check(self.lib.io_net_set_inject(net, feats[0], feats[1], feats[2], self.inject_idx.data_ptr()))
        self.enc_pair_tensor = torch.zeros(self.lib.io_pair_tensor_bytes(self.max_images, self.d), dtype=torch.uint8,
                                           device=self.device)
        self.logits_d = torch.empty((self.max_pairs, 2, 3), dtype=torch.float32, device=self.device)
        self.logits_o = torch.empty((self.max_pairs, 2, 2), dtype=torch.float32, device=self.device)

    def __del__(self):
        try:
            for name in ("enc", "do_net", "oo_net"):
                h = getattr(self, name, None)
                if h:
                    self.lib.io_net_destroy(h)
                    setattr(self, name, None)
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------------------------
    def _load_sub(self, net, sd, sub, in_ch, arch, groups, n_layers, head):
        names, arrs = [], []
        for key, shape in synth.bottleneck_layout(in_ch, arch[0], arch[1], arch[2], groups, n_layers):
            v = None
            for k in synth._sub_key(sub, key):
                if k in sd:
                    v = sd[k]
                    break
            if v is None:
                raise KeyError("InstaDepthNet state_dict: missing '%s'" % synth._sub_key(sub, key)[0])
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            if key == "conv1.weight":
                # 3-channel RGB conv1 -> pair-tensor channels 2..4; 2-channel mask conv1 -> channels 0..1
                full = np.zeros((64, 5, 7, 7), np.float32)
                if in_ch == 3:
                    full[:, 2:5] = a
                else:
                    full[:, 0:2] = a
                a = full
            elif a.ndim == 4 and a.shape[2] == 3 and groups > 1:
                # grouped 3x3 -> dense block-diagonal [cout, cin, 3, 3]: the implicit-GEMM kernels stay dense (the
                # per-image encoder is a few percent of the per-pair trunk work even at 32x redundant FLOPs)
                cout, cg = a.shape[0], a.shape[1]
                og = cout // groups
                full = np.zeros((cout, cg * groups, 3, 3), np.float32)
                for g in range(groups):
                    full[g * og:(g + 1) * og, g * cg:(g + 1) * cg] = a[g * og:(g + 1) * og]
                a = full
            names.append(key.encode())
            arrs.append(a)
        if head is not None:
            for leaf in ("weight", "bias"):
                names.append(("fc." + leaf).encode())
                v = sd[head + "." + leaf]
                arrs.append(np.ascontiguousarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32))
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_numel = (C.c_int64 * n)(*[a.size for a in arrs])
        _lib.check(self.lib.io_net_load_state(net, c_names, c_ptrs, c_numel, n))

    def load_state_dict(self, state_dict):
        """Reference ``InstaDepthNet_od`` state_dict (``pretrained.*``, ``do_net.*``, ``oo_net.*``, ``depth_fc.*``,
        ``occ_fc.*``; ``module.`` prefix optional; ``scratch.*`` / ``pretrained.layer4.*`` are ignored)."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        rx = (synth.RESNEXT_WIDTHS, synth.RESNEXT_OUTS, synth.RESNEXT_BLOCKS)
        self._load_sub(self.enc, sd, "pretrained", 3, rx, synth.RESNEXT_GROUPS, self.enc_layers, None)
        if self.with_disparity:
            self._load_decoder(sd)
        if self.do_net is not None:
            self._load_sub(self.do_net, sd, "do_net", 2, RESNET50, 1, 4, "depth_fc")
        if self.oo_net is not None:
            self._load_sub(self.oo_net, sd, "oo_net", 2, RESNET50, 1, 4, "occ_fc")

    # ---- per batch -----------------------------------------------------------------------------------------------
    def stage_batch(self, items, mode="resize"):
        s, P = super().stage_batch(items, mode)
        jobs = self.resize_jobs
        if len(jobs) > self.max_images:
            raise ValueError("%d images in one batch, engine built for %d (max_images)" % (len(jobs), self.max_images))
        if not hasattr(s, "h_idx"):       # per-slot staging of the two small tables (same life cycle as the slot's buffers)
            s.h_idx = torch.zeros(2 * self.max_pairs, dtype=torch.int32).pin_memory()
            s.d_idx = torch.zeros(2 * self.max_pairs, dtype=torch.int32, device=self.device)
            s.h_enc_desc = torch.zeros(self.max_images * 48, dtype=torch.uint8).pin_memory()
            s.d_enc_desc = torch.zeros(self.max_images * 48, dtype=torch.uint8, device=self.device)
        # one descriptor per image for the encoder's input (its masks are ignored: zero weight columns)
        desc = np.frombuffer(s.h_enc_desc.numpy(), dtype=_lib.PAIR_DESC_DTYPE)
        mask_off = 0
        k = 0
        for (sc, pairs, crops, mat_off, _) in items:
            d = desc[k]
            d["image_off"] = jobs[k][0]
            d["mask_a_off"] = d["mask_b_off"] = mask_off
            d["h"], d["w"], d["x"], d["y"], d["s"], d["rgb_slot"] = sc.h, sc.w, 0, 0, 0, jobs[k][3]
            mask_off += (sc.n * sc.h * sc.w + 15) // 16 * 16
            k += 1
        pdesc = np.frombuffer(s.h_desc.numpy(), dtype=_lib.PAIR_DESC_DTYPE)[:P]
        idx = s.h_idx.numpy()
        idx[0:2 * P:2] = pdesc["rgb_slot"]            # both directions of a pair read their image's encoder features
        idx[1:2 * P:2] = pdesc["rgb_slot"]
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            s.d_enc_desc[:k * 48].copy_(s.h_enc_desc[:k * 48], non_blocking=True)
            s.d_idx[:2 * P].copy_(s.h_idx[:2 * P], non_blocking=True)
            s.copied = torch.cuda.Event()
            s.copied.record(self.copy_stream)
        compute.wait_event(s.copied)
        # the trunks' launch plans hold ONE index pointer: refreshed on the compute stream, behind the previous batch
        self.inject_idx[:2 * P].copy_(s.d_idx[:2 * P], non_blocking=True)
        self.h2d_bytes += k * 48 + 8 * P
        self._last_slot = s
        return s, P

    def upload_resident(self, items, mat_elems, mode="resize"):
        r = super().upload_resident(items, mat_elems, mode)
        r.d_enc_desc = self._last_slot.d_enc_desc.clone()
        r.d_idx = self._last_slot.d_idx.clone()
        return r

    def run_resident(self, r, heads, mode="resize"):
        self.inject_idx[:2 * r.P].copy_(r.d_idx[:2 * r.P], non_blocking=True)
        super().run_resident(r, heads, mode)

    def gather(self, s, P, mode="resize", geom=None):
        if mode != "resize":
            raise NotImplementedError("InstaDepthNet runs in 'resize' mode (its shipped config); got %r" % (mode,))
        super().gather(s, P, mode)
        n_img = len(self.resize_jobs)
        _lib.check(self.lib.io_pair_gather_resize(self._planes.data_ptr(), s.d_mask.data_ptr(),
                                                  s.d_enc_desc.data_ptr(), n_img, self.d,
                                                  self.enc_pair_tensor.data_ptr(), _lib.stream_ptr()))
        self.gpu_launches += 1
        self._n_img = n_img

    def forward(self, P, geom=None):
        if self.do_net is None:
            raise RuntimeError("this engine has no order trunks (plain MiDaS): use disparity() / disparity_order()")
        st = _lib.stream_ptr()
        _lib.check(self.lib.io_net_forward_pairs(self.enc, self.enc_pair_tensor.data_ptr(), self._n_img, None, st))
        _lib.check(self.lib.io_net_forward_pairs(self.do_net, self.pair_tensor.data_ptr(), P, self.logits_d.data_ptr(), st))
        if self.oo_net is not None:
            _lib.check(self.lib.io_net_forward_pairs(self.oo_net, self.pair_tensor.data_ptr(), P, self.logits_o.data_ptr(), st))
            self.logits[:P, :, 0:2] = self.logits_o[:P]
        else:
            self.logits[:P, :, 0:2] = 0      # no occlusion head: sigmoid(0) = 0.5 is never > 0.5, the matrix stays zero
        self.logits[:P, :, 2:5] = self.logits_d[:P]
        self.gpu_launches += sum(self.lib.io_net_last_launches(h) for h in (self.enc, self.do_net, self.oo_net) if h) + 2

    def infer_scenes(self, scenes, algo="InstaDepthNet_od", pairs="all", patch_or_image="resize", return_details=False,
                     _pending=False, _first_cap=None):
        if algo not in ("InstaDepthNet_od", "InstaDepthNet_d"):
            raise ValueError("DepthOrderEngine runs InstaDepthNet_od / _d, got %r" % (algo,))
        return super().infer_scenes(scenes, "InstaOrderNet_od", pairs, patch_or_image, return_details, _pending, _first_cap)

    def submit_scenes(self, scenes, algo="InstaDepthNet_od", pairs="all", patch_or_image="resize", first_batch_pairs=None):
        return self.infer_scenes(scenes, algo, pairs, patch_or_image, False, _pending=True, _first_cap=first_batch_pairs)

    # ---- disparity branch (reference midas_net.py:189-198, midas/blocks.py:124-195) ------------------------------
    def _load_decoder(self, sd):
        dev = self.device
        g = lambda k: (sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else np.asarray(sd[k]))
        dec = {}

        def put(name, wkey, bkey=None, cin_pad=None, cout_pad=None):
            w = _pack_conv(g(wkey), cin_pad, cout_pad).to(dev).contiguous()
            b = np.zeros(w.shape[0], np.float32)
            if bkey is not None:
                bb = g(bkey)
                b[:bb.size] = bb
            dec[name] = (w, torch.from_numpy(b).to(dev))

        for k in range(1, 5):
            put("rn%d" % k, "scratch.layer%d_rn.weight" % k)
            for u in (1, 2):
                for c in (1, 2):
                    p = "scratch.refinenet%d.resConfUnit%d.conv%d" % (k, u, c)
                    put("r%du%dc%d" % (k, u, c), p + ".weight", p + ".bias")
        put("oc0", "scratch.output_conv.0.weight", "scratch.output_conv.0.bias")
        put("oc2", "scratch.output_conv.2.weight", "scratch.output_conv.2.bias", cout_pad=64)      # 32 -> 64 output channels
        put("oc4", "scratch.output_conv.4.weight", "scratch.output_conv.4.bias", cin_pad=64, cout_pad=64)
        self._dec = dec

    def _disparity_current(self):
        """Disparity maps [n_img, D, D] fp32 (CUDA) of the images of the batch the encoder has just processed."""
        if self._dec is None:
            raise RuntimeError("create the engine with with_disparity=True and load a state_dict with scratch.* tensors")
        n, d, dev, st = self._n_img, self.d, self.device, _lib.stream_ptr()
        bf = lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)

        def conv(x_ptr, h, cin, name, cout, k=3, relu=0, res=None):
            w, b = self._dec[name]
            y = bf(n, h, h, cout)
            _lib.check(self.lib.io_conv_bn_act(x_ptr, n, h, h, cin, w.data_ptr(), b.data_ptr(),
                                               res.data_ptr() if res is not None else None, cout, k, 1, relu, y.data_ptr(), st))
            self.gpu_launches += 1
            return y

        def rcu(xp, h, pfx):            # xp = relu(x): blocks.py:146-161 with its in-place ReLU
            t = conv(xp.data_ptr(), h, 256, pfx + "c1", 256, relu=1)
            return conv(t.data_ptr(), h, 256, pfx + "c2", 256, relu=0, res=xp)

        def up(x, h, c, align):
            y = bf(n, 2 * h, 2 * h, c)
            _lib.check(self.lib.io_upsample2x_bilinear(x.data_ptr(), n, h, h, c, align, y.data_ptr(), st))
            self.gpu_launches += 1
            return y

        sizes = [d // 4, d // 8, d // 16, d // 32]
        chans = synth.RESNEXT_OUTS
        rn = [conv(self._enc_feats[k], sizes[k], chans[k], "rn%d" % (k + 1), 256, relu=1) for k in range(4)]   # relu(layer_k_rn)
        path = up(rcu(rn[3], sizes[3], "r4u2"), sizes[3], 256, 1)                                             # refinenet4
        for k in (3, 2, 1):
            h = sizes[k - 1]
            r1 = rcu(rn[k - 1], h, "r%du1" % k)
            o = bf(n, h, h, 256)
            _lib.check(self.lib.io_add_relu(path.data_ptr(), r1.data_ptr(), o.data_ptr(), o.numel(), 1, st))
            self.gpu_launches += 1
            path = up(rcu(o, h, "r%du2" % k), h, 256, 1)
        o = conv(path.data_ptr(), d // 2, 256, "oc0", 128)                     # output_conv.0 (192^2 at D = 384)
        o = up(o, d // 2, 128, 0)                                              # Interpolate(align_corners=False)
        o = conv(o.data_ptr(), d, 128, "oc2", 64, relu=1)                      # output_conv.2 + ReLU (32 real channels)
        o = conv(o.data_ptr(), d, 64, "oc4", 64, k=1, relu=1)                  # output_conv.4 + ReLU (1 real channel)
        return o[..., 0].float()

    def disparity(self, scenes):
        """``InstaDepthNet_od.forward(image, ...)[0]`` for a list of scenes / images -> numpy [len, D, D] fp32."""
        out = []
        for i in range(0, len(scenes), self.max_images):
            chunk = scenes[i:i + self.max_images]
            items = [(sc, np.zeros((0, 2), np.int32), None, 0, k) for k, sc in enumerate(chunk)]
            s, P = self.stage_batch(items, "resize")
            self.gather(s, 0, "resize")
            _lib.check(self.lib.io_net_forward_pairs(self.enc, self.enc_pair_tensor.data_ptr(), self._n_img, None,
                                                     _lib.stream_ptr()))
            self.gpu_launches += self.lib.io_net_last_launches(self.enc)
            out.append(self._disparity_current().cpu().numpy())
            self.finish(s)
        return np.concatenate(out) if out else np.zeros((0, self.d, self.d), np.float32)

    def disparity_order(self, sc, pairs="all", disp_select_method="median"):
        """``infer_order_sup_depth(..., method='InstaDepthNet_d'|'InstaDepthNet_od', disp_select_method='median'|'mean')``
        (reference inference.py:589-599 with ``net_forward_midas_pretrained`` :79-104): pixel depth 1 / (disp + 1e-6), per
        instance clipped to its own 5 % / 95 % quantiles inside the (nearest-resized) modal mask, median or mean; the
        closer instance wins.  The statistic belongs to the INSTANCE, so it is computed N times, not once per pair
        direction.  Returns (int64 [N, N] depth order, disparity clipped to its 5 % / 95 % quantiles as a CPU tensor)."""
        if disp_select_method not in ("median", "mean"):
            raise ValueError("disp_select_method must be 'median' or 'mean'")
        items = [(sc, np.zeros((0, 2), np.int32), None, 0, 0)]
        s, _ = self.stage_batch(items, "resize")
        self.gather(s, 0, "resize")
        _lib.check(self.lib.io_net_forward_pairs(self.enc, self.enc_pair_tensor.data_ptr(), self._n_img, None, _lib.stream_ptr()))
        self.gpu_launches += self.lib.io_net_last_launches(self.enc)
        disp = self._disparity_current()[0]
        d = self.d
        # modal masks at network resolution: cv2.INTER_NEAREST, src = min(floor(dst * (1 / (D / S))), S - 1) (inference.py:398-399)
        mask_bytes = sc.n * sc.h * sc.w
        m = s.d_mask[:mask_bytes].view(sc.n, sc.h, sc.w)
        iy = (torch.arange(d, dtype=torch.float64) * (1.0 / (d / sc.h))).floor().long().clamp(max=sc.h - 1).to(self.device)
        ix = (torch.arange(d, dtype=torch.float64) * (1.0 / (d / sc.w))).floor().long().clamp(max=sc.w - 1).to(self.device)
        mr = m[:, iy][:, :, ix].bool()
        depth = 1.0 / (disp + 1e-6)
        stat = []
        for k in range(sc.n):
            v = depth[mr[k]]
            c = torch.clip(v, torch.quantile(v, 0.05), torch.quantile(v, 0.95))
            stat.append(torch.median(c) if disp_select_method == "median" else torch.mean(c))
        stat = torch.stack(stat).cpu().numpy() if stat else np.zeros(0, np.float32)
        self.finish(s)
        pr = enumerate_pairs(sc.n)
        if pairs == "nbor" and pr.shape[0]:
            pr = pr[self.bordering(sc, pr)]
        order = np.zeros((sc.n, sc.n), np.int64)
        for (i, j) in pr:
            if stat[i] < stat[j]:
                order[i, j], order[j, i] = 1, 0
            elif stat[i] > stat[j]:
                order[i, j], order[j, i] = 0, 1
            else:
                order[i, j] = order[j, i] = 2
        clipped = torch.clip(disp, torch.quantile(disp, 0.05), torch.quantile(disp, 0.95)).cpu()
        self.d2h_bytes += clipped.numel() * 4 + stat.size * 4
        return order, clipped, stat


def _pack_conv(w, cin_pad=None, cout_pad=None):
    """[cout, cin, k, k] fp32 -> bf16 [cout'][k*k*cin'] (tap-major, channel-minor), zero-padded channels."""
    w = np.asarray(w, dtype=np.float32)
    co, ci, k, _ = w.shape
    cip, cop = cin_pad or ci, cout_pad or co
    full = np.zeros((cop, k, k, cip), np.float32)
    full[:co, :, :, :ci] = w.transpose(0, 2, 3, 1)
    return torch.from_numpy(full.reshape(cop, k * k * cip)).to(torch.bfloat16)

def encoder_flops_per_image(d=384):
    """Algorithmic 2*MAC of the encoder's conv1 + layer1..3 (grouped 3x3 counted with their true group size)."""
    total = 2.0 * (d // 2) ** 2 * 64 * 3 * 49
    inpl, side = 64, d // 4
    for li in range(3):
        w, o = synth.RESNEXT_WIDTHS[li], synth.RESNEXT_OUTS[li]
        for b in range(synth.RESNEXT_BLOCKS[li]):
            stride = 2 if (b == 0 and li > 0) else 1
            so = side // stride
            total += 2.0 * side * side * inpl * w                                   # conv1 (before the stride)
            total += 2.0 * so * so * (w // synth.RESNEXT_GROUPS) * 9 * w             # grouped 3x3
            total += 2.0 * so * so * w * o                                           # conv3
            if b == 0:
                total += 2.0 * so * so * inpl * o                                    # downsample
            inpl, side = o, so
    return total
