"""Seeded synthetic scenes and random-init checkpoints.

There is no network and no dataset on the build / GPU boxes, so every test and benchmark runs on synthetic
COCO-shaped scenes: an RGB image, N instance masks (filled ellipses / rectangles) and their xywh boxes in the
format ``datasets/reader.py::InstaOrderDataset.get_image_instances`` (reference ``datasets/reader.py:421-457``)
hands to ``inference.infer_order_sup_*``.  numpy only; no device work happens here.
"""
import numpy as np

COCO_SHAPES = [(480, 640), (427, 640), (640, 480), (375, 500), (333, 500)]


def mask_to_bbox(mask):
    """xywh box of a {0,1} mask; ``[0,0,0,0]`` when empty (reference ``utils/data_utils.py:75-84``)."""
    m = mask == 1
    if not m.any():
        return [0, 0, 0, 0]
    rows = np.where(m.any(axis=1))[0]
    cols = np.where(m.any(axis=0))[0]
    return [int(cols[0]), int(rows[0]), int(cols[-1] + 1 - cols[0]), int(rows[-1] + 1 - rows[0])]


def make_scene(rng, H=480, W=640, N=8, wh_range=((40, 300), (40, 300)), float_boxes=False):
    """One synthetic image: (image u8 [H,W,3], masks u8 [N,H,W] in {0,1}, boxes [N,4] xywh)."""
    image = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    masks = np.zeros((N, H, W), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    boxes = []
    for n in range(N):
        while True:
            w = int(rng.randint(wh_range[0][0], wh_range[0][1] + 1))
            h = int(rng.randint(wh_range[1][0], wh_range[1][1] + 1))
            w, h = min(w, W), min(h, H)
            x = int(rng.randint(0, W - w + 1))
            y = int(rng.randint(0, H - h + 1))
            if rng.rand() < 0.5:
                m = ((xx >= x) & (xx < x + w) & (yy >= y) & (yy < y + h))
            else:
                cx, cy = x + w / 2.0, y + h / 2.0
                m = (((xx + 0.5 - cx) / (w / 2.0)) ** 2 + ((yy + 0.5 - cy) / (h / 2.0)) ** 2) <= 1.0
            if m.any():
                break
        masks[n] = m.astype(np.uint8)
        b = mask_to_bbox(masks[n])
        if float_boxes:  # COCO annotation style: floats with 2 decimals that enclose the mask
            b = [round(b[0] - rng.rand() * 0.5, 2), round(b[1] - rng.rand() * 0.5, 2),
                 round(b[2] + rng.rand(), 2), round(b[3] + rng.rand(), 2)]
        boxes.append(b)
    boxes = np.array(boxes, dtype=np.float64 if float_boxes else np.int64)
    return image, masks, boxes


def make_gt(rng, N):
    """Synthetic GT matrices in the format of ``reader.get_gt_ordering`` (reference ``datasets/reader.py:335-400``):
    occlusion order {0,1} with -1 on the diagonal, depth order {0,1,2} (antisymmetric 0/1, symmetric 2),
    overlap {0,1} symmetric, count {1,2,3} symmetric."""
    occ = (rng.rand(N, N) < 0.25).astype(np.int64)
    np.fill_diagonal(occ, -1)
    depth = np.zeros((N, N), dtype=np.int64)
    overlap = np.zeros((N, N), dtype=np.int64)
    count = np.ones((N, N), dtype=np.int64)
    for i in range(N):
        for j in range(i + 1, N):
            d = int(rng.randint(0, 3))
            depth[i, j] = {0: 1, 1: 0, 2: 2}[d]
            depth[j, i] = {0: 0, 1: 1, 2: 2}[d]
            overlap[i, j] = overlap[j, i] = int(rng.rand() < 0.3)
            count[i, j] = count[j, i] = int(rng.randint(1, 4))
    return occ, depth, overlap, count


def coco_scene_stream(seed, n_images, N=10, ragged=False):
    """C2 workload: COCO-val-shaped images with N instances each (45 pairs at N=10)."""
    rng = np.random.RandomState(seed)
    for _ in range(n_images):
        H, W = COCO_SHAPES[int(rng.randint(0, len(COCO_SHAPES)))]
        n = int(np.clip(rng.poisson(10), 2, 20)) if ragged else N
        yield make_scene(rng, H, W, n, float_boxes=True)


# ----------------------------------------------------------------------------------------------------------------
# ResNet-50 (5-channel, reference models/backbone/resnet_cls.py:119-222) parameter inventory + random init
# ----------------------------------------------------------------------------------------------------------------

def resnet50_layout(in_channels=5, num_classes=(2, 3)):
    """Ordered list of (key, shape) for the reference ``resnet50_cls`` state_dict, *without* the ``module.`` prefix.

    ``num_classes`` int -> ``fc``; list/tuple -> ``fc_occ`` + ``fc_depth`` (reference ``resnet_cls.py:153-160``).
    """
    out = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k, k)))

    def bn(name, c):
        out.append((name + ".weight", (c,)))
        out.append((name + ".bias", (c,)))
        out.append((name + ".running_mean", (c,)))
        out.append((name + ".running_var", (c,)))
        out.append((name + ".num_batches_tracked", ()))

    conv("conv1", 64, in_channels, 7)
    bn("bn1", 64)
    inpl = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for b in range(blocks):
            p = "layer%d.%d" % (li, b)
            conv(p + ".conv1", planes, inpl, 1)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4)
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1)
                bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    if isinstance(num_classes, (list, tuple)):
        out.append(("fc_occ.weight", (num_classes[0], 2048)))
        out.append(("fc_occ.bias", (num_classes[0],)))
        out.append(("fc_depth.weight", (num_classes[1], 2048)))
        out.append(("fc_depth.bias", (num_classes[1],)))
    else:
        out.append(("fc.weight", (num_classes, 2048)))
        out.append(("fc.bias", (num_classes,)))
    return out


def random_state_dict(seed, in_channels=5, num_classes=(2, 3), prefix="module."):
    """Random-init weights of the reference architecture as a dict of numpy fp32 arrays (int64 for
    ``num_batches_tracked``), keyed like a reference checkpoint's ``state_dict``.

    Conv weights are kaiming-normal(fan_out) as in the reference ctor (``resnet_cls.py:162-167``); BN affine and
    running statistics are drawn so that activations keep an O(1) scale through the 16 bottlenecks (the reference's
    own xavier(gain=0.02) re-init gives logits ~1e-12, i.e. all ties -- SURVEY.md fact 5 -- useless for timing *and*
    for parity).  Used by bench.py and smoke(); parity tests use ``oracle/calib.py`` on top of this.
    """
    rng = np.random.RandomState(seed)
    sd = {}
    for key, shape in resnet50_layout(in_channels, num_classes):
        leaf = key.rsplit(".", 1)[1]
        if len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            v = rng.standard_normal(shape).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_out))
        elif leaf == "num_batches_tracked":
            v = np.array(1, dtype=np.int64)
        elif leaf == "running_mean":
            v = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        elif leaf == "running_var":
            v = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif len(shape) == 2:  # fc
            v = (rng.standard_normal(shape) * 0.05).astype(np.float32)
        elif key.startswith("fc"):
            v = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        elif leaf == "weight":  # BN gamma; damp the residual branch so depth does not blow the scale up
            v = rng.uniform(0.2, 0.4, size=shape).astype(np.float32) if ".bn3." in key \
                else rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        else:  # BN beta
            v = (rng.standard_normal(shape) * 0.05).astype(np.float32)
        sd[prefix + key] = v
    return sd


# ----------------------------------------------------------------------------------------------------------------
# InstaDepthNet^od (reference midas/midas_net.py:113-212): ResNeXt-101 32x8d encoder + two 2-channel ResNet-50 trunks.
# Only the tensors the ORDER outputs depend on are generated (encoder layer1-3, do_net, oo_net, depth_fc, occ_fc); the
# MiDaS decoder (scratch.*) and encoder layer4 feed the disparity output only.
# ----------------------------------------------------------------------------------------------------------------
RESNEXT_WIDTHS = (256, 512, 1024, 2048)     # resnext101_32x8d: conv2 width = planes * (8 / 64) * 32, out = planes * 4
RESNEXT_OUTS = (256, 512, 1024, 2048)
RESNEXT_BLOCKS = (3, 4, 23, 3)
RESNEXT_GROUPS = 32


def _sub_key(sub, key):
    """Reference state_dict key of canonical ResNet key ``key`` ('conv1.weight', 'bn1.bias', 'layer2.0.conv1.weight',
    ...) inside sub-module ``sub``: ``pretrained`` wraps conv1/bn1/relu/maxpool/layer1 into ``layer1 = nn.Sequential``
    (midas/blocks.py:_make_resnet_backbone), ``do_net`` / ``oo_net`` do the same but also keep the original attributes
    (midas_net.py:149-161), so conv1 / bn1 / layer1 exist under two names there."""
    head, rest = key.split(".", 1)
    if head == "conv1":
        return ["%s.layer1.0.%s" % (sub, rest)] + (["%s.conv1.%s" % (sub, rest)] if sub != "pretrained" else [])
    if head == "bn1":
        return ["%s.layer1.1.%s" % (sub, rest)] + (["%s.bn1.%s" % (sub, rest)] if sub != "pretrained" else [])
    if head == "layer1":
        return ["%s.layer1.4.%s" % (sub, rest)]
    return ["%s.%s" % (sub, key)]


def bottleneck_layout(in_channels, widths, outs, blocks, groups, n_layers=4):
    out = [("conv1.weight", (64, in_channels, 7, 7))]
    for leaf in ("weight", "bias", "running_mean", "running_var"):
        out.append(("bn1." + leaf, (64,)))
    inpl = 64
    for li in range(n_layers):
        for b in range(blocks[li]):
            p = "layer%d.%d" % (li + 1, b)
            convs = [(".conv1", ".bn1", widths[li], inpl, 1, 1), (".conv2", ".bn2", widths[li], widths[li] // groups, 3, groups),
                     (".conv3", ".bn3", outs[li], widths[li], 1, 1)]
            if b == 0:
                convs.append((".downsample.0", ".downsample.1", outs[li], inpl, 1, 1))
            for (cn, bnn, co, ci, k, _) in convs:
                out.append((p + cn + ".weight", (co, ci, k, k)))
                for leaf in ("weight", "bias", "running_mean", "running_var"):
                    out.append((p + bnn + "." + leaf, (co,)))
            inpl = outs[li]
    return out


def instadepth_state_dict(seed, prefix="module.", with_decoder=False):
    """Random weights for the order branch of InstaDepthNet^od, keyed like the reference model's ``state_dict``
    (numpy fp32).  Same recipe as ``random_state_dict`` (kaiming fan_out convolutions, BN drawn to keep an O(1)
    scale); the parity tests calibrate BN statistics and heads on top (oracle/instadepth_oracle.py).
    ``with_decoder``: also the tensors of the disparity branch (encoder layer4, ``scratch.*``), drawn AFTER the
    order-branch tensors so that those do not depend on the flag."""
    rng = np.random.RandomState(seed)
    sd = {}
    subs = (("pretrained", 3, RESNEXT_WIDTHS, RESNEXT_OUTS, RESNEXT_BLOCKS, RESNEXT_GROUPS, 3),
            ("do_net", 2, (64, 128, 256, 512), (256, 512, 1024, 2048), (3, 4, 6, 3), 1, 4),
            ("oo_net", 2, (64, 128, 256, 512), (256, 512, 1024, 2048), (3, 4, 6, 3), 1, 4))
    for (sub, cin, widths, outs, blocks, groups, n_layers) in subs:
        for key, shape in bottleneck_layout(cin, widths, outs, blocks, groups, n_layers):
            leaf = key.rsplit(".", 1)[1]
            if len(shape) == 4:
                fan_out = shape[0] * shape[2] * shape[3] // (groups if (shape[2] == 3 and groups > 1) else 1)
                v = rng.standard_normal(shape).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_out))
            elif leaf == "running_mean":
                v = (rng.standard_normal(shape) * 0.1).astype(np.float32)
            elif leaf == "running_var":
                v = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
            elif leaf == "weight":
                v = rng.uniform(0.2, 0.4, size=shape).astype(np.float32) if ".bn3." in key \
                    else rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
            else:
                v = (rng.standard_normal(shape) * 0.05).astype(np.float32)
            for k in _sub_key(sub, key):
                sd[prefix + k] = v
    for head, k in (("depth_fc", 3), ("occ_fc", 2)):
        sd[prefix + head + ".weight"] = (rng.standard_normal((k, 2048)) * 0.05).astype(np.float32)
        sd[prefix + head + ".bias"] = (rng.standard_normal(k) * 0.1).astype(np.float32)
    if with_decoder:
        def draw(key, shape, groups=1):
            leaf = key.rsplit(".", 1)[1]
            if len(shape) == 4:
                fan_out = shape[0] * shape[2] * shape[3] // groups
                return rng.standard_normal(shape).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_out))
            if leaf == "running_mean":
                return (rng.standard_normal(shape) * 0.1).astype(np.float32)
            if leaf == "running_var":
                return rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
            if leaf == "weight":
                return rng.uniform(0.2, 0.4, size=shape).astype(np.float32) if ".bn3." in key \
                    else rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
            return (rng.standard_normal(shape) * 0.05).astype(np.float32)
        full = bottleneck_layout(3, RESNEXT_WIDTHS, RESNEXT_OUTS, RESNEXT_BLOCKS, RESNEXT_GROUPS, 4)
        for key, shape in full:
            if key.startswith("layer4."):
                g = RESNEXT_GROUPS if (len(shape) == 4 and shape[2] == 3) else 1
                sd[prefix + "pretrained." + key] = draw(key, shape, g)
        for k, c in enumerate(RESNEXT_OUTS, start=1):
            sd[prefix + "scratch.layer%d_rn.weight" % k] = draw("w.weight", (256, c, 3, 3)) * np.float32(0.5)
        for k in (4, 3, 2, 1):
            for u in (1, 2):
                for c in (1, 2):
                    p = "scratch.refinenet%d.resConfUnit%d.conv%d" % (k, u, c)
                    sd[prefix + p + ".weight"] = draw("w.weight", (256, 256, 3, 3)) * np.float32(0.5)
                    sd[prefix + p + ".bias"] = (rng.standard_normal(256) * 0.05).astype(np.float32)
        for name, shape in (("0", (128, 256, 3, 3)), ("2", (32, 128, 3, 3)), ("4", (1, 32, 1, 1))):
            w = draw("w.weight", shape)
            # the last 1x1 convolution reads post-ReLU features: positive weights keep the (non-negative) disparity
            # away from an all-zero map
            sd[prefix + "scratch.output_conv.%s.weight" % name] = np.abs(w) if name == "4" else w
            sd[prefix + "scratch.output_conv.%s.bias" % name] = (np.abs(rng.standard_normal(shape[0])) * 0.1).astype(np.float32)
    return sd
