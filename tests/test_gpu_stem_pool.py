"""conv1 + bn1 + ReLU + MaxPool2d(3, 2, 1) as one kernel (csrc/stem_pool.cu: parity-split input rows in shared
memory, overlapping no-swizzle UMMA descriptors for the stride-2 im2col, max-pool on the TMEM accumulators) against
torch fp32 of the same ops on the bf16-rounded inputs / weights (reference models/backbone/resnet_cls.py:205-208)."""
import numpy as np
import pytest
import torch

from instaorder_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pairs", [1, 3, 40])
def test_stem_pool_matches_torch(pairs):
    L = _lib.lib()
    D = 256
    g = torch.Generator(device="cuda").manual_seed(pairs)
    rgb = torch.randn((pairs, 3, D, D), generator=g, device="cuda")
    m1 = (torch.rand((pairs, 1, D, D), generator=g, device="cuda") > 0.6).float()
    m2 = (torch.rand((pairs, 1, D, D), generator=g, device="cuda") > 0.6).float()
    pt = torch.zeros(int(L.io_pair_tensor_bytes(pairs, D)), dtype=torch.uint8, device="cuda")
    _lib.check(L.io_pair_pack_nchw(rgb.data_ptr(), m1.data_ptr(), m2.data_ptr(), pairs, D, pt.data_ptr(),
                                   _lib.stream_ptr()))
    w = (torch.randn((64, 5, 7, 7), generator=torch.Generator().manual_seed(7)) * 0.08)
    bias = torch.randn(64, generator=torch.Generator().manual_seed(8)) * 0.3
    out = torch.zeros((2 * pairs, 64, 64, 64), dtype=torch.bfloat16, device="cuda")
    wc, bc = w.contiguous().numpy(), bias.contiguous().numpy()
    _lib.check(L.io_stem_pool(pt.data_ptr(), pairs, D, wc.ctypes.data, bc.ctypes.data, out.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float().cuda()
    xq = torch.cat([m1, m2, rgb], dim=1).to(torch.bfloat16).float()
    for d in range(2):
        x = xq if d == 0 else xq[:, [1, 0, 2, 3, 4]]
        y = torch.nn.functional.conv2d(x, wq, bias.cuda(), stride=2, padding=3)
        y = torch.relu(y).to(torch.bfloat16).float()
        ref = torch.nn.functional.max_pool2d(y, 3, 2, 1).permute(0, 2, 3, 1)          # NHWC
        got = out[d::2].float()
        err = (got - ref).abs()
        tol = 1e-2 + 1e-2 * ref.abs()                                                  # bf16 output rounding
        bad = int((err > tol).sum())
        assert bad == 0, "direction %d: %d elements off, max err %.4f" % (d, bad, float(err.max()))
        # the pool itself is exact: wherever the conv outputs agree the maxima agree, so the typical error is ~0
        assert float(err.mean()) < 2e-3
