"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + torch-CPU fp32) of InstaOrder's pairwise-order hot path.

This file is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``instaorder_b200/`` imports it and the product path never
falls back to it.

Pinning: the reference has no tests or golden vectors of its own, so the oracle is pinned
against the reference *itself* -- ``tests/test_oracle_vs_reference.py`` runs the unmodified reference (through
``oracle/ref_shim.py``) in the build container, and ``oracle/gen_golden.py`` freezes reference outputs into
``tests/golden/*.npz`` for the GPU box.  Third-party arithmetic is pinned to this image's versions: OpenCV 4.13.0
*generic* resize code (IPP HAL disabled: the IPP cubic differs by +-1 u8 LSB on 3-5 % of pixels and is CPU
dependent), torch 2.11.0 fp32 conv/BN, scikit-learn 1.9.0.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
import collections

import numpy as np

DATA_MEAN = [0.485, 0.456, 0.406]   # utils/data_utils.py:9
DATA_STD = [0.229, 0.224, 0.225]    # utils/data_utils.py:10


# ----------------------------------------------------------------------------------------------------------------
# G -- pair construction
# ----------------------------------------------------------------------------------------------------------------

def enumerate_pairs(n):
    """Row-major (i, j), i < j.  inference.py:355-356 / 443-444 / 521-522."""
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def bordering(a, b):
    """inference.py:691-696 -- dilate ``a`` with the 3x3 cross (cv2 border = replicate-of-nothing: out-of-image
    taps are ignored), then any((dilated == 1) & b).  Note ``& b`` is a bitwise AND of a bool with the u8 mask:
    only bit 0 of ``b`` counts."""
    a = a.astype(np.uint8)
    d = a.copy()
    d[1:, :] = np.maximum(d[1:, :], a[:-1, :])
    d[:-1, :] = np.maximum(d[:-1, :], a[1:, :])
    d[:, 1:] = np.maximum(d[:, 1:], a[:, :-1])
    d[:, :-1] = np.maximum(d[:, :-1], a[:, 1:])
    return bool(np.any((d == 1) & (b.astype(np.uint8) & 1).astype(bool)))


def expand_bbox(bboxes, enlarge_box=3.0):
    """tools/test.py:155-163 (inference-only pre-expansion of every instance box)."""
    out = []
    for bbox in bboxes:
        cx = bbox[0] + bbox[2] / 2.
        cy = bbox[1] + bbox[3] / 2.
        size = max([np.sqrt(bbox[2] * bbox[3] * enlarge_box), bbox[2] * 1.1, bbox[3] * 1.1])
        out.append([int(cx - size / 2.), int(cy - size / 2.), int(size), int(size)])
    return np.array(out)


def combine_bbox(bboxes):
    """utils/data_utils.py:61-72."""
    l = bboxes[:, 0].min()
    u = bboxes[:, 1].min()
    r = (bboxes[:, 0] + bboxes[:, 2]).max()
    b = (bboxes[:, 1] + bboxes[:, 3]).max()
    return np.array([l, u, r - l, b - u])


def pair_crop_box(bboxes, i, j):
    """inference.py:361-365 -- square crop window of the pair; ``int()`` truncates toward zero."""
    bbox = combine_bbox(bboxes[(i, j), :])
    cx = bbox[0] + bbox[2] / 2.
    cy = bbox[1] + bbox[3] / 2.
    size = max([np.sqrt(bbox[2] * bbox[3] * 2.), bbox[2] * 1.1, bbox[3] * 1.1])
    return [int(cx - size / 2.), int(cy - size / 2.), int(size), int(size)]


def crop_padding(img, roi, pad_value=0):
    """utils/data_utils.py:104-124 -- w x h window at (x, y); outside the image = pad_value."""
    x, y, w, h = (int(v) for v in roi)
    H, W = img.shape[:2]
    out = np.full((h, w) + img.shape[2:], pad_value, dtype=img.dtype)
    x0, x1, y0, y1 = max(x, 0), min(x + w, W), max(y, 0), min(y + h, H)
    if x1 > x0 and y1 > y0:  # == bbox_iou(...) > 0, utils/data_utils.py:87-101,119
        out[y0 - y:y1 - y, x0 - x:x1 - x] = img[y0:y1, x0:x1]
    return out


def nearest_index(src_len, dst_len):
    """cv2 INTER_NEAREST source index (OpenCV 4.13 resize.cpp resizeNN): floor(dst * (1/(dst_len/src_len))),
    clamped to src_len-1, evaluated in double."""
    inv = np.float64(dst_len) / np.float64(src_len)
    ifx = np.float64(1.0) / inv
    idx = np.floor(np.arange(dst_len, dtype=np.float64) * ifx).astype(np.int64)
    return np.minimum(idx, src_len - 1)


def resize_nearest(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_NEAREST).  inference.py:229-236."""
    ys = nearest_index(src.shape[0], dh)
    xs = nearest_index(src.shape[1], dw)
    return src[ys][:, xs]


def cubic_taps_fixed(src_len, dst_len):
    """cv2 INTER_CUBIC tap table for 8-bit images (OpenCV 4.13 resize.cpp: resize() coefficient loop +
    interpolateCubic, A = -0.75, 11-bit fixed-point coefficients).  Returns (first_tap int32[dst], coef int32[dst,4]);
    tap k reads source index clamp(first_tap + k, 0, src_len-1)."""
    f32 = np.float32
    inv = np.float64(dst_len) / np.float64(src_len)
    scale = np.float64(1.0) / inv
    d = np.arange(dst_len, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int32)
    x = (fx - sx.astype(f32)).astype(f32)
    A = f32(-0.75)
    one = f32(1.0)
    xp1 = (x + one).astype(f32)
    omx = (one - x).astype(f32)
    c0 = ((A * xp1 - f32(5.0) * A) * xp1 + f32(8.0) * A) * xp1 - f32(4.0) * A
    c1 = ((A + f32(2.0)) * x - (A + f32(3.0))) * x * x + one
    c2 = ((A + f32(2.0)) * omx - (A + f32(3.0))) * omx * omx + one
    c3 = one - c0 - c1 - c2
    cb = np.stack([c0, c1, c2, c3], axis=1).astype(f32)
    ic = np.clip(np.rint(cb * f32(2048.0)), -32768, 32767).astype(np.int32)
    return sx - 1, ic


def resize_cubic_u8(src, dw, dh):
    """cv2.resize(src_u8, (dw, dh), interpolation=cv2.INTER_CUBIC) -- bit-exact restatement of OpenCV 4.13's
    generic path (HResizeCubic<uchar,int,short> then VResizeCubicVec_32s8u: fp32 mul/add, *no* fma, taps 3,2,1,0,
    round-half-even, saturate).  inference.py:366-368."""
    f32 = np.float32
    Hs, Ws = src.shape[:2]
    s3 = src.reshape(Hs, Ws, -1)
    tx, ia = cubic_taps_fixed(Ws, dw)
    ty, ib = cubic_taps_fixed(Hs, dh)
    ix = np.clip(tx[:, None] + np.arange(4)[None, :], 0, Ws - 1)
    iy = np.clip(ty[:, None] + np.arange(4)[None, :], 0, Hs - 1)
    h = (s3[:, ix, :].astype(np.int32) * ia[None, :, :, None]).sum(axis=2)            # [Hs, dw, C] int32
    rows = h[iy].astype(f32)                                                          # [dh, 4, dw, C]
    b = (ib.astype(f32) * (f32(1.0) / f32(2048 * 2048))).astype(f32)[:, :, None, None]
    v = (rows[:, 3] * b[:, 3]).astype(f32)
    for k in (2, 1, 0):
        v = ((rows[:, k] * b[:, k]).astype(f32) + v).astype(f32)
    out = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return out.reshape((dh, dw) + src.shape[2:])


def normalize_lut():
    """fp32 value of ``transforms.Normalize(mean, std)(u8 / 255.)`` for every (channel, u8) --
    utils/data_utils.py:28-34 (torchvision: ``t.sub_(mean).div_(std)`` on fp32 tensors)."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    mean = np.array(DATA_MEAN, dtype=np.float32)[:, None]
    std = np.array(DATA_STD, dtype=np.float32)[:, None]
    return ((v[None, :] - mean) / std).astype(np.float32)      # [3, 256]


def transform_rgb(rgb_u8):
    """utils/data_utils.py:28-34 -> fp32 [3, D, D] (without the leading batch dim / .cuda())."""
    lut = normalize_lut()
    return np.stack([lut[c][rgb_u8[:, :, c]] for c in range(3)], axis=0)


def pair_patch(image, inmodal, bboxes, i, j, input_size=256):
    """``patch`` mode of inference.py:360-375: (rgb_u8 [D,D,3], modal_i u8 [D,D], modal_j u8 [D,D], new_bbox)."""
    nb = pair_crop_box(bboxes, i, j)
    if nb[2] <= 0:
        raise ValueError("degenerate pair (%d,%d): crop side int(size) == %d (cv2.resize asserts in the reference)"
                         % (i, j, nb[2]))
    rgb = resize_cubic_u8(crop_padding(image, nb, 0), input_size, input_size)
    mi = resize_nearest(crop_padding(inmodal[i], nb, 0), input_size, input_size)
    mj = resize_nearest(crop_padding(inmodal[j], nb, 0), input_size, input_size)
    return rgb, mi, mj, nb


def pair_tensor(rgb_u8, mi, mj):
    """inference.py:141-145 -- [5, D, D] fp32 in channel order (maskA, maskB, R, G, B)."""
    return np.concatenate([mi[None].astype(np.float32), mj[None].astype(np.float32), transform_rgb(rgb_u8)], axis=0)


def cubic_taps_f64(src_len, dst_len):
    """cv2 INTER_CUBIC taps for CV_64F images (HResizeCubic<double,double,float>): fp32 coefficients, no fixed point."""
    sx, _ = cubic_taps_fixed(src_len, dst_len)
    f32 = np.float32
    inv = np.float64(dst_len) / np.float64(src_len)
    scale = np.float64(1.0) / inv
    d = np.arange(dst_len, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    x = (fx - np.floor(fx).astype(f32)).astype(f32)
    A = f32(-0.75)
    one = f32(1.0)
    xp1 = (x + one).astype(f32)
    omx = (one - x).astype(f32)
    c0 = ((A * xp1 - f32(5.0) * A) * xp1 + f32(8.0) * A) * xp1 - f32(4.0) * A
    c1 = ((A + f32(2.0)) * x - (A + f32(3.0))) * x * x + one
    c2 = ((A + f32(2.0)) * omx - (A + f32(3.0))) * omx * omx + one
    c3 = one - c0 - c1 - c2
    return sx, np.stack([c0, c1, c2, c3], axis=1).astype(f32)


def get_closest_int_multiple_of(orig_num, multiplier):
    """utils/data_utils.py:13-17 (``orig`` mode: ties go up)."""
    r = orig_num % multiplier
    return orig_num + multiplier - r if r >= multiplier // 2 else orig_num - r


def resize_mode_rgb(image, input_size, out_h=None):
    """``resize`` mode rgb: inference.py:395-397 + utils/data_utils.py:37-53 + midas/transforms.py:48-235 --
    image/255. (float64) -> cv2 INTER_CUBIC on CV_64F to input_size^2 (input_size must be a multiple of 32, which
    makes ``Resize(..., ensure_multiple_of=32)`` a plain resize) -> (x-mean)/std in float64 -> CHW fp32.
    ``out_h``: ``orig`` mode (inference.py:401-406), width = input_size, height = out_h, both multiples of 32."""
    out_w, out_h = input_size, (input_size if out_h is None else out_h)
    assert out_w % 32 == 0 and out_h % 32 == 0
    src = image.astype(np.float64) / 255.
    Hs, Ws = src.shape[:2]
    tx, ca = cubic_taps_f64(Ws, out_w)
    ty, cb = cubic_taps_f64(Hs, out_h)
    ix = np.clip(tx[:, None] + np.arange(4)[None, :], 0, Ws - 1)
    iy = np.clip(ty[:, None] + np.arange(4)[None, :], 0, Hs - 1)
    h = (src[:, ix, :] * ca.astype(np.float64)[None, :, :, None]).sum(axis=2)
    out = (h[iy] * cb.astype(np.float64)[:, :, None, None]).sum(axis=1)
    out = (out - np.array(DATA_MEAN)) / np.array(DATA_STD)
    return np.ascontiguousarray(out.transpose(2, 0, 1)).astype(np.float32)


def linear_taps_fixed(src_len, dst_len):
    """cv2 INTER_LINEAR taps for 8-bit images (OpenCV 4.13 resize.cpp): (first index, 11-bit weights a0, a1,
    single-tap flag).  sx < 0 -> (0, fx = 0); sx >= src_len - 1 -> (src_len - 1, fx = 0); when the second tap would
    fall outside, only the first is used with weight 2048."""
    f32 = np.float32
    inv = np.float64(dst_len) / np.float64(src_len)
    scale = np.float64(1.0) / inv
    d = np.arange(dst_len, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx.astype(f32)).astype(f32)
    neg = sx < 0
    fx = np.where(neg, f32(0), fx); sx = np.where(neg, 0, sx)
    hi = sx >= src_len - 1
    fx = np.where(hi, f32(0), fx); sx = np.where(hi, src_len - 1, sx)
    a0 = np.clip(np.rint((f32(1) - fx) * f32(2048)), -32768, 32767).astype(np.int64)
    a1 = np.clip(np.rint(fx * f32(2048)), -32768, 32767).astype(np.int64)
    return sx, a0, a1, (sx + 1 >= src_len)


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src_u8, (dw, dh), interpolation=cv2.INTER_LINEAR), generic path, bit-exact:
    HResizeLinear<uchar,int,short> then VResizeLinear: (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2."""
    f32 = np.float32
    Hs, Ws = src.shape[:2]
    s3 = src.reshape(Hs, Ws, -1).astype(np.int64)
    sx, a0, a1, one = linear_taps_fixed(Ws, dw)
    sx1 = np.minimum(sx + 1, Ws - 1)
    h = np.where(one[None, :, None], s3[:, sx, :] * 2048, s3[:, sx, :] * a0[None, :, None] + s3[:, sx1, :] * a1[None, :, None])
    inv = np.float64(dh) / np.float64(Hs)
    scale = np.float64(1.0) / inv
    fy = ((np.arange(dh, dtype=np.float64) + 0.5) * scale - 0.5).astype(f32)
    sy = np.floor(fy).astype(np.int32)
    fy = (fy - sy.astype(f32)).astype(f32)
    b0 = np.clip(np.rint((f32(1) - fy) * f32(2048)), -32768, 32767).astype(np.int64)
    b1 = np.clip(np.rint(fy * f32(2048)), -32768, 32767).astype(np.int64)
    S0 = h[np.clip(sy, 0, Hs - 1)]
    S1 = h[np.clip(sy + 1, 0, Hs - 1)]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def pad_square(a):
    """inference.py:377-391 -- zero-pad to a centred max(H, W) square."""
    hh, ww = a.shape[:2]
    s = int(max(hh, ww))
    left, top = (s - ww) // 2, (s - hh) // 2
    out = np.zeros((s, s) + a.shape[2:], dtype=a.dtype)
    out[top:top + hh, left:left + ww] = a
    return out


def image_mode_rgb(image, input_size):
    """``image`` mode rgb, inference.py:390-393: padded square -> INTER_LINEAR (u8) -> transform_rgb; CHW fp32."""
    return transform_rgb(resize_linear_u8(pad_square(image), input_size, input_size))


def image_mode_mask(mask, input_size):
    """inference.py:383-388."""
    return resize_nearest(pad_square(mask), input_size, input_size)


def resize_mode_mask(mask, input_size):
    """inference.py:398-399 -- whole-image nearest resize of a modal mask."""
    return resize_nearest(mask, input_size, input_size)


# ----------------------------------------------------------------------------------------------------------------
# N -- network (torch-CPU fp32 reference of the 5-channel ResNet-50; models/backbone/resnet_cls.py:75-222)
# ----------------------------------------------------------------------------------------------------------------

def resnet50_forward(sd, x, prefix="module.", eps=1e-5, bn_override=None, return_features=False):
    """Eval-mode forward from a reference-layout state_dict (torch tensors or numpy).  x: [B,5,H,W] fp32.
    Returns dict of head name -> logits ([B,k] fp32): 'fc' or 'fc_occ' + 'fc_depth'.
    ``bn_override(t, name)`` replaces the eval-mode BN (used by oracle/calib.py to measure batch statistics)."""
    import torch
    import torch.nn.functional as F

    def g(k):
        v = sd[prefix + k]
        return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))

    def bn(t, name):
        if bn_override is not None:
            return bn_override(t, name)
        return F.batch_norm(t, g(name + ".running_mean"), g(name + ".running_var"), g(name + ".weight"),
                            g(name + ".bias"), False, 0.0, eps)

    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        t = F.relu(bn(F.conv2d(x, g("conv1.weight"), stride=2, padding=3), "bn1"))
        t = F.max_pool2d(t, 3, 2, 1)
        for li, blocks in enumerate((3, 4, 6, 3), start=1):
            for b in range(blocks):
                p = "layer%d.%d" % (li, b)
                stride = 2 if (b == 0 and li > 1) else 1
                idt = t
                o = F.relu(bn(F.conv2d(t, g(p + ".conv1.weight")), p + ".bn1"))
                o = F.relu(bn(F.conv2d(o, g(p + ".conv2.weight"), stride=stride, padding=1), p + ".bn2"))
                o = bn(F.conv2d(o, g(p + ".conv3.weight")), p + ".bn3")
                if b == 0:
                    idt = bn(F.conv2d(t, g(p + ".downsample.0.weight"), stride=stride), p + ".downsample.1")
                t = F.relu(o + idt)
        feat = torch.flatten(F.adaptive_avg_pool2d(t, 1), 1)
        out = {}
        if return_features:
            out["features"] = feat.numpy()
        for head in ("fc", "fc_occ", "fc_depth"):
            if (prefix + head + ".weight") in sd:
                out[head] = F.linear(feat, g(head + ".weight"), g(head + ".bias")).numpy()
        return out


def resnet50_forward_bf16(sd, x, prefix="module.", eps=1e-5, round_identity=False):
    """The same network evaluated with *ideal* bf16 storage: BN folded into bf16 weights (fp32 bias), every stored
    activation rounded to bf16, fp32 accumulation -- the arithmetic contract of the CUDA path.  Used to
    separate kernel bugs (CUDA vs this, tight tolerance) from bf16-vs-fp32 sensitivity (this vs resnet50_forward).
    The downsample branch of layerN.0 is accumulated in fp32 together with conv3 (the CUDA path computes both in one
    GEMM and never stores the identity); ``round_identity=True`` emulates the two-launch fallback instead."""
    import torch
    import torch.nn.functional as F

    def r(t):
        return t.to(torch.bfloat16).float()

    def g(k):
        v = sd[prefix + k]
        return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))

    def fold(conv, bn):
        s = (g(bn + ".weight").double() / torch.sqrt(g(bn + ".running_var").double() + eps))
        w = r((g(conv + ".weight") * s.float()[:, None, None, None]))
        b = (g(bn + ".bias").double() - g(bn + ".running_mean").double() * s).float()
        return w, b[None, :, None, None]

    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        w, b = fold("conv1", "bn1")
        t = r(F.relu(F.conv2d(r(x), w, stride=2, padding=3) + b))
        t = F.max_pool2d(t, 3, 2, 1)
        for li, blocks in enumerate((3, 4, 6, 3), start=1):
            for blk in range(blocks):
                p = "layer%d.%d" % (li, blk)
                stride = 2 if (blk == 0 and li > 1) else 1
                idt = t
                w, b = fold(p + ".conv1", p + ".bn1")
                o = r(F.relu(F.conv2d(t, w) + b))
                w, b = fold(p + ".conv2", p + ".bn2")
                o = r(F.relu(F.conv2d(o, w, stride=stride, padding=1) + b))
                w, b = fold(p + ".conv3", p + ".bn3")
                o = F.conv2d(o, w) + b
                if blk == 0:
                    w, b = fold(p + ".downsample.0", p + ".downsample.1")
                    idt = F.conv2d(t, w, stride=stride) + b
                    if round_identity:
                        idt = r(idt)
                t = r(F.relu(o + idt))
        feat = torch.flatten(F.adaptive_avg_pool2d(t, 1), 1)
        out = {}
        for head in ("fc", "fc_occ", "fc_depth"):
            if (prefix + head + ".weight") in sd:
                out[head] = F.linear(feat, g(head + ".weight"), g(head + ".bias")).numpy()
        return out


# ----------------------------------------------------------------------------------------------------------------
# H -- heads / decisions / order matrices
# ----------------------------------------------------------------------------------------------------------------

def _softmax(z):
    z = z.astype(np.float32)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


def _sigmoid(z):
    return (1.0 / (1.0 + np.exp(-z.astype(np.float32)))).astype(np.float32)


def decide_occ(logit1, logit2):
    """inference.py:196-214 (H1).  logit1 = f(A,B), logit2 = f(B,A), each [2].
    Returns (A_over_B, B_over_A, margin) with margin = min |p - 0.5|."""
    o1, o2 = _sigmoid(logit1), _sigmoid(logit2)
    p12 = (o1[1] + o2[0]) / np.float32(2)
    p21 = (o1[0] + o2[1]) / np.float32(2)
    return bool(p12 > 0.5), bool(p21 > 0.5), float(min(abs(p12 - 0.5), abs(p21 - 0.5)))


def decide_depth(logit1, logit2):
    """inference.py:172-193 (H2).  Returns (argidx in {0 closer, 1 farther, 2 equal}, margin = top1 - top2)."""
    d1, d2 = _softmax(logit1), _softmax(logit2)
    p = np.array([(d1[0] + d2[1]) / 2, (d1[1] + d2[0]) / 2, (d1[2] + d2[2]) / 2], dtype=np.float32)
    s = np.sort(p)
    return int(np.argmax(p)), float(s[-1] - s[-2])


def decide_ordernet(logit1, logit2):
    """inference.py:44-76 (H4), 3- or 4-way.  Returns (A_over_B, B_over_A, margin)."""
    o1, o2 = _softmax(logit1), _softmax(logit2)
    p = [(o1[1] + o2[0]) / 2, (o1[0] + o2[1]) / 2, (o1[2] + o2[2]) / 2,
         (o1[3] + o2[3]) / 2 if o1.shape[-1] == 4 else np.float32(0)]
    p = np.array(p, dtype=np.float32)
    s = np.sort(p)
    a = int(np.argmax(p))
    return (a in (0, 3)), (a in (1, 3)), float(s[-1] - s[-2])


def write_depth(mat, i, j, argidx):
    """inference.py:416-428 / 612-623."""
    if argidx == 0:
        mat[i, j], mat[j, i] = 1, 0
    elif argidx == 1:
        mat[i, j], mat[j, i] = 0, 1
    else:
        mat[i, j], mat[j, i] = 2, 2


def write_occ(mat, i, j, i_over_j, j_over_i):
    """inference.py:430-434 / 507-510."""
    if i_over_j:
        mat[i, j] = 1
    if j_over_i:
        mat[j, i] = 1


def infer_order(sd, image, inmodal, bboxes, pairs="all", method="InstaOrderNet_od", patch_or_image="patch",
                input_size=256, forward=None, chunk=8, batch1=False):
    """The public drivers inference.py:349-436 / 439-512 / 515-624 (H6) restated with a batched fp32 forward
    (``batch1=True``: one forward per network input, i.e. two batch-1 forwards per pair exactly as the reference's
    ``net_forward_*`` do -- the faithful CPU baseline of bench.py when the reference archive is absent).

    Returns dict(occ=int64[N,N] | None, depth=int64[N,N] | None, margin_occ, margin_depth, logits=...)."""
    forward = forward or resnet50_forward
    N = inmodal.shape[0]
    plist = [(i, j) for (i, j) in enumerate_pairs(N) if pairs == "all" or bordering(inmodal[i], inmodal[j])]
    occ = np.zeros((N, N), dtype=np.int64)
    depth = np.zeros((N, N), dtype=np.int64)
    m_occ = np.full((N, N), np.inf)
    m_depth = np.full((N, N), np.inf)
    rgb_whole = resize_mode_rgb(image, input_size) if patch_or_image == "resize" else \
        (image_mode_rgb(image, input_size) if patch_or_image == "image" else None)
    logits = {}
    for c0 in range(0, len(plist), chunk):
        xs = []
        for (i, j) in plist[c0:c0 + chunk]:
            if patch_or_image == "patch":
                rgb, mi, mj, _ = pair_patch(image, inmodal, bboxes, i, j, input_size)
                x = pair_tensor(rgb, mi, mj)
            elif patch_or_image == "resize":
                mi = resize_mode_mask(inmodal[i], input_size).astype(np.float32)
                mj = resize_mode_mask(inmodal[j], input_size).astype(np.float32)
                x = np.concatenate([mi[None], mj[None], rgb_whole], axis=0)
            elif patch_or_image == "image":
                mi = image_mode_mask(inmodal[i], input_size).astype(np.float32)
                mj = image_mode_mask(inmodal[j], input_size).astype(np.float32)
                x = np.concatenate([mi[None], mj[None], rgb_whole], axis=0)
            elif patch_or_image == "orig":       # inference.py:401-408: H, W rounded to multiples of 32, non-square network input
                hh = get_closest_int_multiple_of(inmodal.shape[1], 32)
                ww = get_closest_int_multiple_of(inmodal.shape[2], 32)
                if rgb_whole is None:
                    rgb_whole = resize_mode_rgb(image, ww, hh)
                mi = resize_nearest(inmodal[i], ww, hh).astype(np.float32)
                mj = resize_nearest(inmodal[j], ww, hh).astype(np.float32)
                x = np.concatenate([mi[None], mj[None], rgb_whole], axis=0)
            else:
                raise NotImplementedError(patch_or_image)
            xs.append(x)
            xs.append(x[[1, 0, 2, 3, 4]])
        if batch1:
            outs = [forward(sd, x[None].astype(np.float32)) for x in xs]
            out = {h: np.concatenate([o[h] for o in outs]) for h in outs[0]}
        else:
            out = forward(sd, np.stack(xs).astype(np.float32))
        for k, (i, j) in enumerate(plist[c0:c0 + chunk]):
            lg = {h: (v[2 * k], v[2 * k + 1]) for h, v in out.items()}
            logits[(i, j)] = lg
            if method == "InstaOrderNet_od":
                a, b, m = decide_occ(*lg["fc_occ"])
                write_occ(occ, i, j, a, b)
                m_occ[i, j] = m_occ[j, i] = m
                d, m = decide_depth(*lg["fc_depth"])
                write_depth(depth, i, j, d)
                m_depth[i, j] = m_depth[j, i] = m
            elif method == "InstaOrderNet_o":
                a, b, m = decide_occ(*lg["fc"])
                write_occ(occ, i, j, a, b)
                m_occ[i, j] = m_occ[j, i] = m
            elif method == "OrderNet":
                a, b, m = decide_ordernet(*lg["fc"])
                write_occ(occ, i, j, a, b)
                m_occ[i, j] = m_occ[j, i] = m
            elif method == "InstaOrderNet_d":
                d, m = decide_depth(*lg["fc"])
                write_depth(depth, i, j, d)
                m_depth[i, j] = m_depth[j, i] = m
            else:
                raise ValueError(method)
    return dict(occ=occ, depth=depth, margin_occ=m_occ, margin_depth=m_depth, logits=logits, pairs=plist)


def infer_gt_order(inmodal, amodal):
    """inference.py:719-739 (M4)."""
    n = inmodal.shape[0]
    gt = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for j in range(i + 1, n):
            if not bordering(inmodal[i], inmodal[j]):
                continue
            oij = int(((inmodal[i] == 1) & (amodal[j] == 1)).sum())
            oji = int(((inmodal[j] == 1) & (amodal[i] == 1)).sum())
            if oij == 0 and oji == 0:
                continue
            if oij >= oji:
                gt[i, j], gt[j, i] = 1, 0
            else:
                gt[i, j], gt[j, i] = 0, 1
    return gt


# ----------------------------------------------------------------------------------------------------------------
# M -- metrics
# ----------------------------------------------------------------------------------------------------------------

def eval_order_recall_precision_f1(order, gt, zd):
    """inference.py:794-802 -- sklearn binary recall/precision/F1 over all entries (diagonal included) with
    gt != -1, x100; each score -> float(zd) when its denominator is 0.  Returns (recall, precision, f1).
    (Positive class is label 1; a pred of 2 is neither TP nor FP, as in sklearn's binary average.)"""
    keep = gt != -1
    g = gt[keep].reshape(-1)
    p = order[keep].reshape(-1)
    tp = int(np.sum((g == 1) & (p == 1)))
    fp = int(np.sum((g != 1) & (p == 1)))
    fn = int(np.sum((g == 1) & (p != 1)))
    r = tp / (tp + fn) if (tp + fn) > 0 else float(zd)
    pr = tp / (tp + fp) if (tp + fp) > 0 else float(zd)
    f = 2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else float(zd)
    return r * 100, pr * 100, f * 100


WHDR_KEYS = ["%s_%s" % (o, e) for o in ("ovlX", "ovlO", "ovlOX") for e in ("eq", "neq", "all")]


def eval_depth_order_whdr(order, gt_order_ovl_count):
    """inference.py:764-791 (+ calculate_whdr :757-761, extract_upper_tri_without_diagonal :17-19).
    Returns defaultdict(list) with the 9 keys '{ovl}_{eq}', each a 1-element list (float or -1)."""
    gt, ovl, cnt = gt_order_ovl_count
    iu = np.triu_indices_from(gt, k=1)
    gt, ovl, cnt, pred = gt[iu], ovl[iu], cnt[iu], order[iu]
    score = 2 / cnt
    m_ovl = collections.OrderedDict()
    m_ovl["ovlX"] = ovl == 0
    m_ovl["ovlO"] = ovl == 1
    m_ovl["ovlOX"] = m_ovl["ovlX"] | m_ovl["ovlO"]
    m_eq = collections.OrderedDict()
    m_eq["eq"] = gt == 2
    m_eq["neq"] = (gt == 0) | (gt == 1)
    m_eq["all"] = m_eq["eq"] | m_eq["neq"]
    out = collections.defaultdict(list)
    for ko, mo in m_ovl.items():
        for ke, me in m_eq.items():
            m = mo & me
            if m.sum() == 0:
                out["%s_%s" % (ko, ke)].append(-1)
            else:
                out["%s_%s" % (ko, ke)].append(((gt[m] != pred[m]) * score[m]).sum() / score[m].sum() * 100)
    return out
