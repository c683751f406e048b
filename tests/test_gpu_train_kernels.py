"""Per-kernel parity of the training-step building blocks against torch fp32 autograd of the same op on the same
bf16-rounded operands (floating point; tolerances stated per test): tcgen05 weight gradient (MN-major operands),
data gradient (forward weights read MN-major, mirrored taps), train-mode BatchNorm forward / backward, max-pool with
arg-max, fused SGD / Adam."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from instaorder_b200 import _lib
import gpu_util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


# (B, H, W, Cin, Cout, k, stride): every layer family of the ResNet-50 at 256^2 and at the 64^2 test size
WGRAD_CASES = [
    (2, 64, 64, 64, 64, 1, 1),      # layer1.0 conv1: GEMM mode, cout 64 (one A slab)
    (2, 64, 64, 256, 64, 1, 1),     # layer1 conv1: N tile 256
    (2, 64, 64, 64, 256, 1, 1),     # layer1 conv3 / downsample: two M tiles
    (2, 64, 64, 64, 64, 3, 1),      # layer1 conv2: 3x3, one row of 64 pixels per K block
    (3, 32, 32, 128, 128, 3, 1),    # layer2 conv2: 2 rows x 32
    (3, 16, 16, 256, 256, 3, 1),    # layer3 conv2: 4 rows x 16
    (5, 8, 8, 512, 512, 3, 1),      # layer4 conv2: one image per K block, odd batch, 2 N tiles
    (2, 64, 64, 128, 128, 3, 2),    # layer2.0 conv2: stride 2
    (3, 16, 16, 512, 512, 3, 2),    # layer4.0 conv2: stride 2, 8x8 out
    (2, 64, 64, 256, 512, 1, 2),    # layer2.0 downsample: 1x1 stride 2
    (6, 8, 8, 2048, 512, 1, 1),     # layer4 conv1: 8 N tiles
    (6, 8, 8, 512, 2048, 1, 1),     # layer4 conv3: 16 M tiles
    (5, 4, 4, 256, 256, 3, 1),      # 64^2 net layer3: 4 images per K block, ragged batch (OOB images)
    (7, 2, 2, 512, 512, 3, 1),      # 64^2 net layer4: 16 images per K block
    (4, 4, 4, 512, 512, 3, 2),      # 64^2 net layer4.0 conv2: stride 2 to 2x2
    (37, 8, 8, 512, 2048, 1, 1),    # rows not a multiple of 64
    (2, 24, 24, 256, 256, 3, 1),    # 384^2 geometry: K block = 2 rows x 24 = 48 pixels
]


def _ids(c):
    return "B%d_%dx%d_%d-%d_k%ds%d" % c


@pytest.mark.parametrize("case", WGRAD_CASES, ids=_ids)
def test_conv_wgrad(case):
    """dW vs torch.nn.grad.conv2d_weight in fp32 on the bf16-rounded x, dy.  Tolerance: fp32 accumulation in a
    different order over up to 10^5 terms -> 2e-3 of the per-tensor gradient scale."""
    B, H, W, Cin, Cout, k, stride = case
    g = _gen(sum(case))
    Ho, Wo = H // stride, W // stride
    x = torch.randn((B, H, W, Cin), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    dy = torch.randn((B, Ho, Wo, Cout), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    dw = torch.zeros((Cout, k * k * Cin), device=DEV, dtype=torch.float32)
    _lib.check(_lib.lib().io_conv_wgrad(x.data_ptr(), B, H, W, Cin, dy.data_ptr(), Cout, k, stride, dw.data_ptr(),
                                        _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, k, k),
                                      dy.float().permute(0, 3, 1, 2), stride=stride, padding=k // 2)
    ref = ref.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)
    scale = float(ref.abs().max())
    err = float((dw - ref).abs().max())
    assert torch.isfinite(dw).all()
    assert err <= 2e-3 * scale, "max |dW - ref| = %.4g (scale %.4g)" % (err, scale)


DGRAD_CASES = [
    (2, 64, 64, 64, 64, 1), (2, 64, 64, 256, 64, 1), (2, 64, 64, 64, 256, 1), (2, 64, 64, 64, 64, 3),
    (3, 32, 32, 128, 128, 3), (3, 16, 16, 256, 256, 3), (5, 8, 8, 512, 512, 3), (6, 8, 8, 2048, 512, 1),
    (6, 8, 8, 512, 2048, 1), (5, 4, 4, 256, 256, 3), (7, 2, 2, 512, 512, 3), (2, 32, 32, 256, 512, 1),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "B%d_%dx%d_%d-%d_k%d" % c)
@pytest.mark.parametrize("use_res", [False, True], ids=["", "res"])
def test_conv_dgrad(case, use_res):
    """dx (bf16) vs the fp32 autograd data gradient; tolerance = bf16 output rounding (1e-2 relative + 1e-2)."""
    B, H, W, Cin, Cout, k = case
    g = _gen(sum(case) + 7)
    w = (torch.randn((Cout, Cin, k, k), generator=g, device=DEV) / (Cout * k * k) ** 0.5).to(torch.bfloat16).float()
    dy = torch.randn((B, H, W, Cout), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    res = torch.randn((B, H, W, Cin), generator=g, device=DEV).to(torch.bfloat16).contiguous() if use_res else None
    dx = torch.full((B, H, W, Cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    zero = torch.zeros(2048, device=DEV)
    wp = U.pack_weight(w)
    _lib.check(_lib.lib().io_conv_dgrad(dy.data_ptr(), B, H, W, Cin, Cout, k, wp.data_ptr(), zero.data_ptr(),
                                        res.data_ptr() if use_res else None, dx.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w, dy.float().permute(0, 3, 1, 2), stride=1, padding=k // 2)
    ref = ref.permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.float()
    got = dx.float()
    assert torch.isfinite(got).all(), "%d unwritten outputs" % int((~torch.isfinite(got)).sum())
    err = (got - ref).abs()
    tol = 1e-2 + 1e-2 * ref.abs()
    assert not (err > tol).any(), "max err %.4g, %d bad" % (float(err.max()), int((err > tol).sum()))


# the data gradient through the CTA-pair kernel (N = Cin >= 256 tiles): MN-major weight slabs split between the two CTAs
@pytest.mark.parametrize("case", [c for c in DGRAD_CASES if c[3] >= 256], ids=lambda c: "pair_B%d_%dx%d_%d-%d_k%d" % c)
@pytest.mark.parametrize("use_res", [False, True], ids=["", "res"])
def test_conv_dgrad_pair_kernel(case, use_res, monkeypatch):
    monkeypatch.setenv("INSTAORDER_PAIR", "2")
    test_conv_dgrad(case, use_res)


@pytest.mark.parametrize("pairs,d", [(3, 256), (2, 64), (2, 128)])
def test_stem_wgrad(pairs, d):
    """two-direction stem weight gradient from the pair tensor vs fp32 autograd of the 5-channel 7x7 s2 conv."""
    L = _lib.lib()
    g = _gen(pairs + d)
    rgb = torch.randn((pairs, 3, d, d), generator=g, device=DEV)
    m1 = (torch.rand((pairs, 1, d, d), generator=g, device=DEV) > 0.6).float()
    m2 = (torch.rand((pairs, 1, d, d), generator=g, device=DEV) > 0.6).float()
    pt = torch.zeros(L.io_pair_tensor_bytes(pairs, d), dtype=torch.uint8, device=DEV)
    _lib.check(L.io_pair_pack_nchw(rgb.data_ptr(), m1.data_ptr(), m2.data_ptr(), pairs, d, pt.data_ptr(),
                                   _lib.stream_ptr()))
    ho = d // 2
    dy = torch.randn((2, pairs, ho, ho, 64), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    scratch = torch.zeros((128, 448), device=DEV)
    _lib.check(L.io_stem_wgrad(pt.data_ptr(), pairs, d, dy.data_ptr(), scratch.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    xb = torch.cat([m1, m2, rgb], 1).to(torch.bfloat16).float()
    for direction in range(2):
        x = xb if direction == 0 else xb[:, [1, 0, 2, 3, 4]]
        ref = torch.nn.grad.conv2d_weight(x, (64, 5, 7, 7), dy[direction].float().permute(0, 3, 1, 2), stride=2,
                                          padding=3)          # [64, 5, 7, 7]
        got = scratch[direction * 64:(direction + 1) * 64].reshape(64, 7, 8, 8)[:, :, :7, :5].permute(0, 3, 1, 2)
        if direction == 1:   # rows 64.. are gradients w.r.t. the pair tensor's channels: (B, A) swaps channels 0 / 1
            got = got[:, [1, 0, 2, 3, 4]]
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        assert err <= 2e-3 * scale, "direction %d: max err %.4g (scale %.4g)" % (direction, err, scale)
    # the padding channels of the packed layout only ever see zero inputs (the 8th tap sees real pixels: its forward
    # weight is zero and its gradient is discarded by the unpack kernel)
    pad = scratch.reshape(128, 7, 8, 8)
    assert float(pad[:, :, :, 5:].abs().max()) == 0.0


BN_CASES = [(2, 4096 * 3, 64, True, False), (2, 1024 * 2, 256, True, True), (2, 640, 2048, True, True),
            (2, 200, 512, False, False), (1, 333, 128, True, False), (2, 64 * 64 * 4, 64, True, False)]


@pytest.mark.parametrize("case", BN_CASES, ids=lambda c: "g%d_r%d_c%d%s%s" % (c[0], c[1], c[2], "_relu" if c[3] else "",
                                                                             "_res" if c[4] else ""))
def test_bn_train_forward_backward(case):
    """train-mode BN (+ residual) (+ ReLU) per group vs torch F.batch_norm(training=True) autograd in fp32.
    Tolerances: activations bf16 rounding (1e-2 + 1e-2 |ref|); running stats 1e-4; dgamma / dbeta 2e-3 of scale;
    dy bf16 rounding."""
    groups, rows, c, relu, use_res = case
    L = _lib.lib()
    g = _gen(rows + c)
    y = (torch.randn((groups, rows, c), generator=g, device=DEV) * 1.5 + 0.3).to(torch.bfloat16).contiguous()
    res = torch.randn((groups, rows, c), generator=g, device=DEV).to(torch.bfloat16).contiguous() if use_res else None
    gamma = torch.rand(c, generator=g, device=DEV) + 0.5
    beta = torch.randn(c, generator=g, device=DEV) * 0.2
    rm = torch.randn(c, generator=g, device=DEV) * 0.1
    rv = torch.rand(c, generator=g, device=DEV) + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    a = torch.full((groups, rows, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    save = torch.zeros(4 * groups * c, device=DEV)
    bits = torch.zeros(groups * rows * c // 8, device=DEV, dtype=torch.uint8)
    scratch = torch.zeros(groups * 2 * c, device=DEV, dtype=torch.float64)
    _lib.check(L.io_bn_train_forward(y.data_ptr(), res.data_ptr() if use_res else None, a.data_ptr(), groups, rows, c,
                                     gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(),
                                     save.data_ptr(), scratch.data_ptr(), int(relu),
                                     bits.data_ptr() if (relu and use_res) else None, _lib.stream_ptr()))
    # ReLU mask: read from the stored activation when a residual was added, recomputed from y otherwise
    mask_mode = 0 if not relu else (3 if use_res else 2)
    da = torch.randn((groups, rows, c), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    dy = torch.full((groups, rows, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    gout = torch.full((groups, rows, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    dgamma = torch.zeros(c, device=DEV)
    dbeta = torch.zeros(c, device=DEV)
    _lib.check(L.io_bn_train_backward(da.data_ptr(), bits.data_ptr() if mask_mode == 3 else None, y.data_ptr(),
                                      dy.data_ptr(), gout.data_ptr(),
                                      groups, rows, c, gamma.data_ptr(), save.data_ptr(), scratch.data_ptr(),
                                      mask_mode, dgamma.data_ptr(), dbeta.data_ptr(), _lib.stream_ptr()))
    if mask_mode == 3:   # the mask read from the stored activation (mode 1) must give the same result as the bit mask
        dy1 = torch.empty_like(dy)
        dg1, db1 = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        _lib.check(L.io_bn_train_backward(da.data_ptr(), a.data_ptr(), y.data_ptr(), dy1.data_ptr(), None, groups, rows,
                                          c, gamma.data_ptr(), save.data_ptr(), scratch.data_ptr(), 1, dg1.data_ptr(),
                                          db1.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(dy1, dy)
    torch.cuda.synchronize()
    # reference: one F.batch_norm call per group (the reference's two forward passes), shared gamma / beta
    yf = y.float().requires_grad_(True)
    gm = gamma.clone().requires_grad_(True)
    bt = beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm0.clone(), rv0.clone()
    outs = []
    for q in range(groups):
        o = F.batch_norm(yf[q].t().reshape(1, c, rows), rm_ref, rv_ref, gm, bt, True, 0.1, 1e-5)
        o = o.reshape(c, rows).t()
        if use_res:
            o = o + res[q].float()
        if relu:
            o = torch.relu(o)
        outs.append(o)
    out = torch.stack(outs)
    err = (a.float() - out).abs()
    assert not (err > 1e-2 + 1e-2 * out.abs()).any(), "forward max err %.4g" % float(err.max())
    assert float((rm - rm_ref).abs().max()) < 1e-4 and float((rv - rv_ref).abs().max()) < 1e-4
    # backward on the kernel's own (bf16) activation mask so that ReLU ties do not matter
    mask = (a.float() > 0).float() if relu else torch.ones_like(out)
    pre = []
    for q in range(groups):
        o = F.batch_norm(yf[q].t().reshape(1, c, rows), None, None, gm, bt, True, 0.1, 1e-5).reshape(c, rows).t()
        pre.append(o)
    pre = torch.stack(pre)
    (pre * (da.float() * mask)).sum().backward()
    sc = float(gm.grad.abs().max())
    assert float((dgamma - gm.grad).abs().max()) <= 2e-3 * sc + 1e-3, "dgamma %.4g" % float((dgamma - gm.grad).abs().max())
    sc = float(bt.grad.abs().max())
    assert float((dbeta - bt.grad).abs().max()) <= 2e-3 * sc + 1e-3, "dbeta %.4g" % float((dbeta - bt.grad).abs().max())
    err = (dy.float() - yf.grad).abs()
    assert not (err > 1e-2 + 1e-2 * yf.grad.abs()).any(), "dy max err %.4g" % float(err.max())
    assert torch.equal(gout.float(), (da.float() * mask).to(torch.bfloat16).float())


@pytest.mark.parametrize("b,h,w", [(2, 128, 128), (3, 32, 32), (1, 8, 6)])
def test_maxpool_train(b, h, w):
    """MaxPool2d(3, 2, 1) forward (bit-exact on bf16 values) and backward vs torch autograd (inputs made tie-free)."""
    L = _lib.lib()
    g = _gen(b + h + w)
    c = 64
    # post-ReLU-like input with exact zeros; ties are broken like ATen does (first maximum in window scan order)
    x = torch.relu(torch.randn((b, h, w, c), generator=g, device=DEV)).to(torch.bfloat16).contiguous()
    y = torch.empty((b, h // 2, w // 2, c), device=DEV, dtype=torch.bfloat16)
    idx = torch.empty((b, h // 2, w // 2, c), device=DEV, dtype=torch.uint8)
    dy = torch.randn((b, h // 2, w // 2, c), generator=g, device=DEV).to(torch.bfloat16).contiguous()
    dx = torch.full((b, h, w, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    _lib.check(L.io_maxpool_train(x.data_ptr(), y.data_ptr(), idx.data_ptr(), dy.data_ptr(), dx.data_ptr(), b, h, w, c,
                                  _lib.stream_ptr()))
    torch.cuda.synchronize()
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xf, 3, 2, 1)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    ref.backward(dy.float().permute(0, 3, 1, 2))
    want = xf.grad.permute(0, 2, 3, 1)
    got = dx.float()
    err = (got - want.to(torch.bfloat16).float()).abs()
    frac_bad = float((err > 1e-2 + 1e-2 * want.abs()).float().mean())
    assert frac_bad < 1e-3, "fraction of mismatching gradient entries %.4g" % frac_bad


def test_optim_sgd_adam():
    """fused SGD (momentum, weight decay) and Adam vs torch.optim on the same flat buffers: fp32, 1e-6 relative."""
    L = _lib.lib()
    g = _gen(5)
    n = 100003
    w0 = torch.randn(n, generator=g, device=DEV)
    grads = [torch.randn(n, generator=g, device=DEV) * 0.1 for _ in range(3)]
    # SGD
    w = w0.clone(); buf = torch.zeros(n, device=DEV); w16 = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.SGD([p], lr=1e-2, momentum=0.9, weight_decay=1e-4)
    for it, gr in enumerate(grads):
        _lib.check(L.io_optim_sgd(w.data_ptr(), gr.data_ptr(), buf.data_ptr(), n, 1e-2, 0.9, 1e-4, int(it == 0),
                                  w16.data_ptr(), n, _lib.stream_ptr()))
        p.grad = gr.clone(); opt.step()
    torch.cuda.synchronize()
    assert float((w - p.data).abs().max()) < 1e-6
    assert torch.equal(w16, w.to(torch.bfloat16))
    # Adam
    w = w0.clone(); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.5, 0.999))
    for it, gr in enumerate(grads):
        _lib.check(L.io_optim_adam(w.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.5, 0.999, 1e-8,
                                   it + 1, None, 0, _lib.stream_ptr()))
        p.grad = gr.clone(); opt.step()
    torch.cuda.synchronize()
    assert float((w - p.data).abs().max()) < 2e-6
