"""Validation surface of the model wrappers (reference trainer.py:218-266 -> set_input + forward_only) and
``model.model(x)`` on collated batches, against fixtures produced by the reference's own wrappers."""
import os

import numpy as np
import pytest
import torch

from instaorder_b200 import _lib, models
from oracle import calib, gen_golden

pytestmark = pytest.mark.gpu


def make_model(case):
    c = gen_golden.CASES[case]
    params = dict(algo=c["algo"], backbone_arch="resnet50_cls",
                  backbone_param=dict(in_channels=5, num_classes=c["num_classes"]), optim="SGD", lr=1e-4,
                  weight_decay=1e-4, use_rgb=True, overlap_weight=1.5, distinct_weight=0.5, max_pairs=4)
    m = models.__dict__[c["algo"]](params, dist_model=False)
    m.load_state_dict(calib.load_calibrated(gen_golden.calib_path(case), c["wseed"], 5, c["num_classes"]))
    m.switch_to("eval")
    return m, c


@pytest.mark.parametrize("case", sorted(gen_golden.VAL_CASES))
def test_forward_only_matches_reference(golden_dir, case):
    z = np.load(os.path.join(golden_dir, "val_losses.npz"))
    model, c = make_model(case)
    batch = gen_golden.val_batch(gen_golden.VAL_CASES[case], c["algo"], c["num_classes"])
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    # model.model(x): logits of the (A,B) direction, tolerance 2e-2 (bf16 vs the reference's fp32)
    x = torch.cat([tb["modal1"], tb["modal2"], tb["rgb"]], dim=1).cuda()
    y = model.model(x)
    y = torch.cat(y, dim=1) if isinstance(y, tuple) else y
    ref = z[case + "_logits"]
    assert np.abs(y.cpu().numpy() - ref[:, 0]).max() < 2e-2
    # loss kernel alone on the reference's own fp32 logits: arithmetic identical up to fp32 summation order
    k_total = ref.shape[2]
    lg = torch.from_numpy(ref).cuda().contiguous()
    out = torch.empty(3, dtype=torch.float32, device="cuda")
    algo = c["algo"]
    occ_t = tb["occ_order"].float().cuda().contiguous() if algo in ("InstaOrderNet_od", "InstaOrderNet_o") else None
    cls_t = (tb["depth_order"] if "depth_order" in tb else tb["occ_order"]).long().cuda().contiguous() \
        if algo != "InstaOrderNet_o" else None
    ovl = tb["is_overlap"].long().cuda().contiguous() if algo == "InstaOrderNet_od" else None
    occ_off = 0 if occ_t is not None else -1
    cls_off = {"InstaOrderNet_od": 2, "InstaOrderNet_d": 0, "OrderNet": 0, "InstaOrderNet_o": -1}[algo]
    cls_k = 0 if cls_off < 0 else (k_total - cls_off)
    _lib.check(_lib.lib().io_loss_forward(lg.data_ptr(), lg.shape[0], k_total, occ_off, cls_off, cls_k,
                                          _lib.ptr(occ_t), _lib.ptr(cls_t), _lib.ptr(ovl), 1.5, 0.5, 1,
                                          out.data_ptr(), _lib.stream_ptr()))
    got = out.cpu().numpy()
    assert abs(got[0] - z[case + "_loss"]) < 2e-6 * max(1.0, abs(z[case + "_loss"])) + 1e-6, (got, z[case + "_loss"])
    if algo == "InstaOrderNet_od":
        assert abs(got[1] - z[case + "_loss_occ"]) < 3e-6 and abs(got[2] - z[case + "_loss_depth"]) < 3e-6
    # the wrapper end to end: set_input + forward_only, same structure as the reference's return value
    model.set_input(**tb)
    r = model.forward_only()
    log, loss = r
    assert abs(float(loss["loss"]) - float(z[case + "_loss"])) < 2e-2
    if algo == "InstaOrderNet_od":
        assert set(log) == {"loss_occ", "loss_depth"}
        assert abs(float(log["loss_occ"]) - float(z[case + "_loss_occ"])) < 2e-2
        assert abs(float(log["loss_depth"]) - float(z[case + "_loss_depth"])) < 2e-2
    else:
        assert log == {}
        assert model.forward_only(ret_loss=False) == {}
    # swapped-direction labels of set_input (reference :38-48 etc.)
    if hasattr(model, "depth_order2"):
        d1 = model.depth_order1.cpu().numpy(); d2 = model.depth_order2.cpu().numpy()
        assert np.array_equal(d2, np.where(d1 == 0, 1, np.where(d1 == 1, 0, d1)))
    if algo in ("InstaOrderNet_od", "InstaOrderNet_o"):
        assert torch.equal(model.occ_order2, model.occ_order1[:, [1, 0]])


def test_save_state_round_trip(tmp_path):
    model, c = make_model("c2_od")
    model.save_state(str(tmp_path), 7)
    ck = torch.load(os.path.join(str(tmp_path), "ckpt_iter_7.pth.tar"), map_location="cpu", weights_only=False)
    assert ck["step"] == 7 and all(k.startswith("module.") for k in ck["state_dict"])
    m2 = models.InstaOrderNet_od(model.params)
    assert m2.load_state(str(tmp_path), Iter=7) == 7
    with pytest.raises(Exception):
        m2.load_state(str(tmp_path), Iter=8)          # "=> no checkpoint found" as in the reference
