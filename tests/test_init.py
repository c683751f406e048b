"""N2: a freshly constructed model holds the reference constructor's initialisation --
kaiming, then ``init_weights('xavier', 0.02)`` -- drawn from torch's global CPU generator in the reference's order, so
that after the same ``torch.manual_seed`` the ``state_dict`` is bit-identical (reference
models/single_stage_model.py:24-25, utils/common_utils.py:35-65, models/backbone/resnet_cls.py:162-167)."""
import numpy as np
import pytest
import torch

from instaorder_b200 import init as I
from oracle import ref_shim

CASES = [("InstaOrderNet_od", [2, 3]), ("InstaOrderNet_o", 2), ("InstaOrderNet_d", 3), ("OrderNet", 4)]


def _params(algo, nc):
    return dict(algo=algo, backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=nc),
                optim="SGD", lr=1e-4, weight_decay=1e-4, use_rgb=True)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
@pytest.mark.parametrize("algo,nc", CASES[:2] + CASES[3:])
def test_fresh_state_dict_equals_reference_constructor(algo, nc):
    ns = ref_shim.load()
    torch.manual_seed(7)
    want = ns.models.__dict__[algo](_params(algo, nc), dist_model=False).model.state_dict()
    after_ref = torch.rand(1).item()                     # the generator must also end in the same state
    torch.manual_seed(7)
    got = I.reference_init_state_dict(nc)
    after = torch.rand(1).item()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k
    assert after == after_ref


def test_fresh_state_dict_statistics():
    """Runs without the reference: layout (322 entries for ^od), xavier-normal gain 0.02
    statistics, BatchNorm N(1, 0.02), zero biases, untouched running statistics."""
    torch.manual_seed(0)
    sd = I.reference_init_state_dict([2, 3])
    assert len(sd) == 322
    assert sum(v.numel() for k, v in sd.items() if k.endswith((".weight", ".bias"))) == 23524549
    w = sd["module.layer3.2.conv2.weight"]
    assert tuple(w.shape) == (256, 256, 3, 3)
    want_std = 0.02 * np.sqrt(2.0 / (256 * 9 + 256 * 9))
    assert abs(float(w.std()) / want_std - 1) < 0.02
    g = sd["module.layer4.0.bn3.weight"]
    assert abs(float(g.mean()) - 1) < 2e-3 and abs(float(g.std()) / 0.02 - 1) < 0.1
    assert float(sd["module.fc_depth.bias"].abs().max()) == 0 and float(sd["module.bn1.bias"].abs().max()) == 0
    assert float(sd["module.bn1.running_var"].min()) == 1 and int(sd["module.bn1.num_batches_tracked"]) == 0
    torch.manual_seed(0)
    again = I.reference_init_state_dict([2, 3])
    assert all(torch.equal(sd[k], again[k]) for k in sd)
