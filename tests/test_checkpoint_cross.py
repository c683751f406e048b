"""Cross-implementation checkpoint round trips (reference
models/single_stage_model.py:54-72, utils/common_utils.py:128-149).

* CPU, needs the reference tree: the UNMODIFIED reference trains two steps (real ``torch.optim`` state) and writes
  ``ckpt_iter_2.pth.tar`` with its own ``save_state``; this package's wrapper loads it with ``load_state(resume=True)``
  and writes it back with ``save_state``; a second reference model loads THAT file with its own
  ``load_state(resume=True)`` -- weights, BN buffers and the optimiser state are bit-identical at both ends.
* GPU: the same file layout goes through the flat device buffers -- a checkpoint written with real
  ``torch.optim.SGD`` / ``Adam`` state (the reference's layout) is imported into ``TrainEngine`` / ``FlatOptim``,
  exported again and loaded into a fresh ``torch.optim`` optimiser exactly as ``utils.load_state`` does."""
import os

import numpy as np
import pytest
import torch

from instaorder_b200 import init as I
from instaorder_b200 import models
from oracle import ref_shim


def _params(algo, nc, optim):
    return dict(algo=algo, backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=nc),
                optim=optim, lr=1e-2, weight_decay=1e-4, beta1=0.9, use_rgb=True, overlap_weight=0.1,
                distinct_weight=0.9)


def _same_optim_state(a, b):
    assert a["param_groups"][0]["params"] == b["param_groups"][0]["params"]
    assert set(a["state"].keys()) == set(b["state"].keys())
    for i in a["state"]:
        for k, v in a["state"][i].items():
            w = b["state"][i][k]
            if torch.is_tensor(v):
                assert torch.equal(v.cpu().float(), torch.as_tensor(w).cpu().float()), (i, k)
            else:
                assert v == w, (i, k)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
@pytest.mark.parametrize("optim", ["SGD", "Adam"])
def test_reference_written_checkpoint_round_trip(optim, tmp_path):
    import torch.distributed as dist
    from oracle import gen_golden_train as G
    from oracle import train_oracle as T
    ns = ref_shim.load()
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29573")
        dist.init_process_group("gloo", rank=0, world_size=1)
    algo, nc = "InstaOrderNet_od", [2, 3]
    torch.manual_seed(3)
    ref = ns.models.__dict__[algo](_params(algo, nc, optim), dist_model=False)
    ref.switch_to("train")
    for it in range(2):
        ref.set_input(**G.set_input_args(algo, T.make_batch(40 + it, 2, 64, algo)))
        ref.step()
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(), d2.mkdir()
    ref.save_state(str(d1), 2)

    ours = models.InstaOrderNet_od(_params(algo, nc, optim))
    assert ours.load_state(str(d1), Iter=2, resume=True) == 2
    ours.save_state(str(d2), 2)
    raw = torch.load(str(d2 / "ckpt_iter_2.pth.tar"), map_location="cpu", weights_only=False)
    assert set(raw.keys()) == {"step", "state_dict", "optimizer"} and raw["step"] == 2

    torch.manual_seed(4)
    ref2 = ns.models.__dict__[algo](_params(algo, nc, optim), dist_model=False)
    ref2.load_state(str(d2), Iter=2, resume=True)
    a, b = ref.model.state_dict(), ref2.model.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    _same_optim_state(ref.optim.state_dict(), ref2.optim.state_dict())
    # the resumed reference optimiser takes the same next step as the original
    batch = T.make_batch(50, 2, 64, algo)
    for m in (ref, ref2):
        m.switch_to("train")
        m.set_input(**G.set_input_args(algo, batch))
        m.step()
    for k in a:
        assert torch.equal(ref.model.state_dict()[k], ref2.model.state_dict()[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("optim", ["SGD", "Adam"])
def test_torch_optim_state_through_flat_buffers(optim, tmp_path):
    """A checkpoint in the reference's layout carrying REAL torch.optim state -> load_state(resume=True) -> flat device
    buffers -> save_state -> a fresh torch optimiser's load_state_dict (utils/common_utils.py:143-147)."""
    nc = [2, 3]
    torch.manual_seed(11)
    sd = I.reference_init_state_dict(nc)
    names = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    plist = [torch.nn.Parameter(sd[k].clone()) for k in names]

    def make_opt(ps):
        return torch.optim.SGD(ps, lr=1e-2, momentum=0.9, weight_decay=1e-4) if optim == "SGD" else \
            torch.optim.Adam(ps, lr=1e-2, betas=(0.9, 0.999))
    opt = make_opt(plist)
    g = torch.Generator().manual_seed(5)
    for _ in range(2):
        for p in plist:
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        opt.step()
    state = {k: (plist[names.index(k)].detach().clone() if k in names else v) for k, v in sd.items()}
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(), d2.mkdir()
    torch.save({"step": 2, "state_dict": state, "optimizer": opt.state_dict()}, str(d1 / "ckpt_iter_2.pth.tar"))

    m = models.InstaOrderNet_od(_params("InstaOrderNet_od", nc, optim))
    assert m.load_state(str(d1), Iter=2, resume=True) == 2
    m._train_engine(2, 64)                  # binds the optimiser state to the flat device buffers
    m._train_dirty = True                   # force the export of the device copies
    m.save_state(str(d2), 2)
    ck = torch.load(str(d2 / "ckpt_iter_2.pth.tar"), map_location="cpu", weights_only=False)
    assert list(ck["state_dict"].keys()) == list(state.keys())
    for k in state:
        assert torch.equal(ck["state_dict"][k].float(), state[k].float()), k
    plist2 = [torch.nn.Parameter(torch.zeros_like(p)) for p in plist]
    opt2 = make_opt(plist2)
    opt2.load_state_dict(ck["optimizer"])
    _same_optim_state(opt.state_dict(), opt2.state_dict())
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
