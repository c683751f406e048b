"""Host-side logic of the InstaDepthNet engine that needs no GPU: the dense block-diagonal expansion of grouped 3x3
weights (what ``DepthOrderEngine._load_sub`` passes to ``io_net_load_state``) is the same convolution, the reference
key mapping covers the order branch, and the algorithmic FLOP count used by bench.py."""
import numpy as np
import torch
import torch.nn.functional as F

from instaorder_b200 import depth_engine, synth


def expand(a, groups):
    cout, cg = a.shape[0], a.shape[1]
    og = cout // groups
    full = np.zeros((cout, cg * groups, 3, 3), np.float32)
    for g in range(groups):
        full[g * og:(g + 1) * og, g * cg:(g + 1) * cg] = a[g * og:(g + 1) * og]
    return full


def test_block_diagonal_expansion_equals_grouped_conv():
    rng = np.random.RandomState(0)
    w = rng.standard_normal((64, 2, 3, 3)).astype(np.float32)      # 32 groups of 2 -> 2 channels
    x = torch.from_numpy(rng.standard_normal((2, 64, 9, 11)).astype(np.float32))
    a = F.conv2d(x, torch.from_numpy(w), padding=1, stride=2, groups=32)
    b = F.conv2d(x, torch.from_numpy(expand(w, 32)), padding=1, stride=2)
    assert torch.allclose(a, b, atol=1e-5)


def test_state_dict_layout_and_flops():
    keys = [k for k, _ in synth.bottleneck_layout(3, synth.RESNEXT_WIDTHS, synth.RESNEXT_OUTS, synth.RESNEXT_BLOCKS,
                                                  synth.RESNEXT_GROUPS, 3)]
    assert keys[0] == "conv1.weight" and "layer3.22.conv3.weight" in keys and not any(k.startswith("layer4") for k in keys)
    assert synth._sub_key("pretrained", "conv1.weight") == ["pretrained.layer1.0.weight"]
    assert synth._sub_key("do_net", "bn1.bias") == ["do_net.layer1.1.bias", "do_net.bn1.bias"]
    assert synth._sub_key("oo_net", "layer1.2.conv3.weight") == ["oo_net.layer1.4.2.conv3.weight"]
    assert synth._sub_key("pretrained", "layer3.5.bn2.running_var") == ["pretrained.layer3.5.bn2.running_var"]
    gf = depth_engine.encoder_flops_per_image(384) / 1e9
    assert 80 < gf < 95       # ~86 GFLOP: conv1 + ResNeXt-101 32x8d layer1-3 at 384^2
