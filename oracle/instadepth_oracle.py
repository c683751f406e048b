"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the ORDER outputs of the reference's InstaDepthNet^od
(``midas/midas_net.py:113-212``; BASELINE config 5): the ResNeXt-101 32x8d encoder's
layer1..3 (``pretrained.layer1-3``, :186-188), the two 2-channel ResNet-50 trunks ``do_net`` / ``oo_net`` with the
encoder features added in front of their layer2 / layer3 / layer4 (:200-210) and the heads ``depth_fc`` / ``occ_fc``.
The order matrices of ``infer_order_sup_occ_depth(method="InstaDepthNet_od")`` (``inference.py:349-436, 107-137``)
do not depend on the disparity output; ``disparity_forward`` restates that branch too (encoder layer4, ``scratch.*`` =
the MiDaS decoder, :189-198, ``midas/blocks.py:124-195``) -- groundwork for the CUDA path of the disparity map, which
is not built yet.

Pinned against the unmodified reference: ``oracle/gen_golden_instadepth.py`` builds the reference module (torch.hub
patched to torchvision's architecture-identical ``resnext101_32x8d``, SURVEY.md 8c shim 6), loads the synthetic
calibrated weights into it and freezes its logits in ``tests/golden/instadepth_order.npz``;
``tests/test_instadepth_oracle.py`` checks this file against them."""
import numpy as np

from instaorder_b200 import synth

EPS = 1e-5


def _g(sd, prefix, sub, key):
    import torch
    v = sd[prefix + synth._sub_key(sub, key)[0]]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _trunk(sd, prefix, sub, x, blocks, groups, n_layers, inject, bn_override, rnd, fold):
    """conv1/bn1/relu/maxpool + ``n_layers`` bottleneck layers (torchvision / resnet_cls.py Bottleneck, stride on the
    3x3).  ``inject[l]`` (broadcast over the batch by ``inject_index``) is added to the output of layer l+1 before the
    next layer, as midas_net.py:201-203 does.  Returns the list of (possibly injected) layer outputs."""
    import torch
    import torch.nn.functional as F

    def conv_bn(t, conv, bn, stride=1, padding=0, g=1, relu=True, add=None):
        w = _g(sd, prefix, sub, conv + ".weight")
        if fold:   # the CUDA path's arithmetic: BN folded into bf16 weights + fp32 bias, bf16 storage
            s = (_g(sd, prefix, sub, bn + ".weight").double() /
                 torch.sqrt(_g(sd, prefix, sub, bn + ".running_var").double() + EPS))
            b = (_g(sd, prefix, sub, bn + ".bias").double() - _g(sd, prefix, sub, bn + ".running_mean").double() * s).float()
            y = F.conv2d(t, rnd(w * s.float()[:, None, None, None]), stride=stride, padding=padding, groups=g) + \
                b[None, :, None, None]
        else:
            y = F.conv2d(t, w, stride=stride, padding=padding, groups=g)
            if bn_override is not None:
                y = bn_override(y, sub + "." + bn)
            else:
                y = F.batch_norm(y, _g(sd, prefix, sub, bn + ".running_mean"), _g(sd, prefix, sub, bn + ".running_var"),
                                 _g(sd, prefix, sub, bn + ".weight"), _g(sd, prefix, sub, bn + ".bias"), False, 0.0, EPS)
        if add is not None:
            y = y + add
        return F.relu(y) if relu else y

    t = rnd(conv_bn(rnd(x), "conv1", "bn1", stride=2, padding=3))
    t = F.max_pool2d(t, 3, 2, 1)
    outs = []
    for li in range(n_layers):
        for b in range(blocks[li]):
            p = "layer%d.%d" % (li + 1, b)
            stride = 2 if (b == 0 and li > 0) else 1
            o = rnd(conv_bn(t, p + ".conv1", p + ".bn1"))
            o = rnd(conv_bn(o, p + ".conv2", p + ".bn2", stride=stride, padding=1, g=groups))
            if b == 0:   # identity kept in fp32: the CUDA path computes conv3 + downsample as one GEMM
                idt = conv_bn(t, p + ".downsample.0", p + ".downsample.1", stride=stride, relu=False)
            else:
                idt = t
            t = rnd(conv_bn(o, p + ".conv3", p + ".bn3", add=idt))
        if inject is not None and li < len(inject):
            t = rnd(t + inject[li])
        outs.append(t)
    return outs


def order_forward(sd, rgb, m1, m2, img_index=None, prefix="module.", bn_override=None, bf16=False,
                  return_features=False):
    """rgb [I,3,H,W] (one entry per IMAGE), m1 / m2 [B,1,H,W] (one entry per pair direction), img_index [B] = image of
    each entry (default: entry b uses image b).  Returns dict(depth=[B,3], occ=[B,2]) fp32 logits
    (+ 'feat_depth', 'feat_occ', 'enc' when ``return_features``).  ``bf16``: evaluate with the CUDA path's storage
    rounding (BN folded into bf16 weights, every stored activation bf16, fp32 accumulation)."""
    import torch
    import torch.nn.functional as F
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.float32))
    rnd = (lambda a: a.to(torch.bfloat16).float()) if bf16 else (lambda a: a)
    with torch.no_grad():
        rgb, m1, m2 = t(rgb), t(m1), t(m2)
        idx = torch.arange(m1.shape[0]) if img_index is None else torch.as_tensor(np.asarray(img_index), dtype=torch.long)
        enc = _trunk(sd, prefix, "pretrained", rgb, synth.RESNEXT_BLOCKS, synth.RESNEXT_GROUPS, 3, None, bn_override,
                     rnd, bf16)                                                       # midas_net.py:186-188
        inject = [e[idx] for e in enc]
        x = torch.cat([m1, m2], dim=1)                                                # :200, :207
        out = {}
        for sub, head, name in (("do_net", "depth_fc", "depth"), ("oo_net", "occ_fc", "occ")):
            f4 = _trunk(sd, prefix, sub, x, (3, 4, 6, 3), 1, 4, inject, bn_override, rnd, bf16)[3]     # :200-203 / :207-210
            feat = torch.flatten(F.adaptive_avg_pool2d(f4, 1), 1)                     # :205-206
            w, b = sd[prefix + head + ".weight"], sd[prefix + head + ".bias"]
            out[name] = F.linear(feat, t(w), t(b)).numpy()
            if return_features:
                out["feat_" + name] = feat.numpy()
        if return_features:
            out["enc"] = [e.numpy() for e in enc]
        return out


def disparity_forward(sd, rgb, prefix="module.", bf16=False):
    """``InstaDepthNet_od.forward(...)[0]`` (midas_net.py:186-198): rgb [I,3,H,W] -> disparity [I,H,W] fp32.
    Notes on the reference's arithmetic: ``ResidualConvUnit`` uses an in-place ReLU on its input, so its skip adds
    ``relu(x)``, not ``x`` (blocks.py:146-161); the fusion blocks up-sample with ``align_corners=True`` (:191-193) but
    ``output_conv``'s ``Interpolate`` with ``align_corners=False`` (:117-119).
    ``bf16``: the CUDA path's storage rounding (bf16 weights and stored activations, fp32 accumulation): with these
    random weights the decoder turns bf16 rounding into a ~4 % (of the range) shift of the map, so kernel correctness
    is judged against this emulation and only loosely against the fp32 fixture."""
    import torch
    import torch.nn.functional as F
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a, dtype=np.float32))
    g = lambda k: t(sd[prefix + k])
    r = (lambda a: a.to(torch.bfloat16).float()) if bf16 else (lambda a: a)
    with torch.no_grad():
        enc = _trunk(sd, prefix, "pretrained", t(rgb), synth.RESNEXT_BLOCKS, synth.RESNEXT_GROUPS, 4, None, None, r, bf16)

        def conv(x, name, bias=True, relu=False, res=None, pad=1):
            y = F.conv2d(x, r(g(name + ".weight")), g(name + ".bias") if bias else None, padding=pad)
            if res is not None:
                y = y + res
            return r(F.relu(y) if relu else y)

        def rcu(xp, p):        # xp = relu(x) (in-place ReLU of the reference): conv2(relu(conv1(xp))) + xp
            return conv(conv(xp, p + ".conv1", relu=True), p + ".conv2", res=xp)

        up = lambda x, a: r(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=a))
        rn = [conv(e, "scratch.layer%d_rn" % (k + 1), bias=False, relu=True) for k, e in enumerate(enc)]
        path = up(rcu(rn[3], "scratch.refinenet4.resConfUnit2"), True)
        for k in (3, 2, 1):
            r1 = rcu(rn[k - 1], "scratch.refinenet%d.resConfUnit1" % k)
            o = r(F.relu(path + r1))
            path = up(rcu(o, "scratch.refinenet%d.resConfUnit2" % k), True)
        o = conv(path, "scratch.output_conv.0")
        o = up(o, False)
        o = conv(o, "scratch.output_conv.2", relu=True)
        o = conv(o, "scratch.output_conv.4", relu=True, pad=0)
        return torch.squeeze(o, dim=1).numpy()


# ---- calibrated synthetic checkpoint (same idea as oracle/calib.py) -----------------------------------------------
def calib_inputs(seed=321, n_scenes=4, D=384):
    from oracle import oracle as O
    rng = np.random.RandomState(seed)
    rgbs, m1, m2, idx = [], [], [], []
    for s in range(n_scenes):
        H, W = synth.COCO_SHAPES[int(rng.randint(0, len(synth.COCO_SHAPES)))]
        image, masks, _ = synth.make_scene(rng, H, W, 3)
        rgbs.append(O.resize_mode_rgb(image, D))
        mm = [O.resize_mode_mask(m, D)[None].astype(np.float32) for m in masks]
        for (i, j) in O.enumerate_pairs(3):
            for (a, b) in ((i, j), (j, i)):
                m1.append(mm[a]); m2.append(mm[b]); idx.append(s)
    return np.stack(rgbs).astype(np.float32), np.stack(m1), np.stack(m2), np.asarray(idx)


def calibrate(sd, prefix="module.", logit_std=0.5, seed=321, D=384):
    """In place: BN running statistics measured on synthetic inputs (every BN of the encoder's layer1-3 and of the two
    trunks), heads rebuilt along the principal directions of the pooled features.  Returns the changed tensors."""
    import torch
    rgb, m1, m2, idx = calib_inputs(seed, D=D)
    changed = {}

    def bn_override(y, name):          # name = '<sub>.<canonical bn key>'
        sub, key = name.split(".", 1)
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        rm = (mean * 0.9).numpy().astype(np.float32)
        rv = (var * 1.1 + 1e-3).numpy().astype(np.float32)
        for k in synth._sub_key(sub, key + ".running_mean"):
            sd[prefix + k] = rm; changed[prefix + k] = rm
        for k in synth._sub_key(sub, key + ".running_var"):
            sd[prefix + k] = rv; changed[prefix + k] = rv
        w = torch.from_numpy(sd[prefix + synth._sub_key(sub, key + ".weight")[0]])
        b = torch.from_numpy(sd[prefix + synth._sub_key(sub, key + ".bias")[0]])
        z = (y - torch.from_numpy(rm)[None, :, None, None]) / torch.sqrt(torch.from_numpy(rv) + EPS)[None, :, None, None]
        return z * w[None, :, None, None] + b[None, :, None, None]

    out = order_forward(sd, rgb, m1, m2, idx, prefix, bn_override=bn_override, return_features=True)
    rng = np.random.RandomState(seed + 1)
    for name, head in (("depth", "depth_fc"), ("occ", "occ_fc")):
        feat = out["feat_" + name].astype(np.float64)
        mu = feat.mean(axis=0)
        _, _, vt = np.linalg.svd(feat - mu, full_matrices=False)
        k = sd[prefix + head + ".weight"].shape[0]
        r = min(6, vt.shape[0])
        w = rng.standard_normal((k, r)) @ vt[:r]
        s = float(((feat - mu) @ w.T).std()) + 1e-12
        w = (w * (logit_std / s)).astype(np.float32)
        b = (-(w.astype(np.float64) @ mu) + rng.standard_normal(k) * 0.1 * logit_std).astype(np.float32)
        sd[prefix + head + ".weight"] = w; changed[prefix + head + ".weight"] = w
        sd[prefix + head + ".bias"] = b; changed[prefix + head + ".bias"] = b
    return changed


def load_calibrated(npz_path, seed, prefix="module.", with_decoder=False):
    sd = synth.instadepth_state_dict(seed, prefix, with_decoder)
    with np.load(npz_path) as z:
        for k in z.files:
            assert k in sd and sd[k].shape == z[k].shape, k
            sd[k] = z[k]
    return sd
