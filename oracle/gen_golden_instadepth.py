"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/instadepth_{calib,order}.npz by running the UNMODIFIED
reference (``midas/midas_net.py`` InstaDepthNet_od through ``inference.infer_order_sup_occ_depth`` with
``method="InstaDepthNet_od"``, ``patch_or_image="resize"``, 384^2) on CPU in this container.

    python -m oracle.gen_golden_instadepth

torch.hub is patched to return torchvision's resnext101_32x8d (same architecture as the WSL hub model, no network);
the synthetic calibrated weights (instaorder_b200.synth.instadepth_state_dict + oracle.instadepth_oracle.calibrate)
are loaded into the reference module; tensors the order outputs do not depend on (encoder layer4, the MiDaS decoder,
the trunks' unused fc) keep the reference's own initialisation."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import synth  # noqa: E402
from oracle import instadepth_oracle as IO, ref_shim  # noqa: E402

SEED, SCENE_SEED, N_INST, D = 11, 4242, 4, 384
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_scene():
    rng = np.random.RandomState(SCENE_SEED)
    return synth.make_scene(rng, 375, 500, N_INST, wh_range=((60, 260), (60, 220)))


def main():
    import torch
    import torchvision
    torch.hub.load = lambda *a, **k: torchvision.models.resnext101_32x8d(weights=None)
    ns = ref_shim.load()
    midas_net = sys.modules["_instaorder_ref.midas.midas_net"]
    sd = synth.instadepth_state_dict(SEED, prefix="")
    changed = IO.calibrate(sd, prefix="")
    np.savez_compressed(os.path.join(GOLDEN, "instadepth_calib.npz"), **{"module." + k: v for k, v in changed.items()})
    torch.manual_seed(0)
    net = midas_net.InstaDepthNet_od(path=None)
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("scratch.", "pretrained.layer4", "do_net.fc", "oo_net.fc")) or "num_batches_tracked" in k
               for k in missing), [k for k in missing][:5]
    net.eval()
    image, masks, boxes = build_scene()
    logits = {}

    class Rec(torch.nn.Module):          # records the raw logits of every forward the reference makes
        def __init__(self, m):
            super().__init__()
            self.m = m
            self.calls = []

        def forward(self, img, a, b):
            disp, d, o = self.m(img, a, b)
            self.calls.append((d.numpy().copy(), o.numpy().copy()))
            return disp, d, o

    rec = Rec(net)
    model = types.SimpleNamespace(model=rec)
    occ, depth = ns.inference.infer_order_sup_occ_depth(model, image, masks, boxes, "all", "InstaDepthNet_od", "resize",
                                                        D, "")
    P = N_INST * (N_INST - 1) // 2
    assert len(rec.calls) == 2 * P
    dl = np.stack([c[0][0] for c in rec.calls]).reshape(P, 2, 3)
    ol = np.stack([c[1][0] for c in rec.calls]).reshape(P, 2, 2)
    np.savez_compressed(os.path.join(GOLDEN, "instadepth_order.npz"), occ=occ.astype(np.int64), depth=depth.astype(np.int64),
                        depth_logits=dl.astype(np.float32), occ_logits=ol.astype(np.float32))
    print("occ\n", occ, "\ndepth\n", depth, "\nlogit std", dl.std(), ol.std())
    # the restatement against what the reference just computed
    from oracle import oracle as O
    rgb = O.resize_mode_rgb(image, D)[None]
    mm = [O.resize_mode_mask(m, D)[None, None].astype(np.float32) for m in masks]
    m1, m2 = [], []
    for (i, j) in O.enumerate_pairs(N_INST):
        m1 += [mm[i][0], mm[j][0]]
        m2 += [mm[j][0], mm[i][0]]
    sdm = {"module." + k: v for k, v in sd.items()}
    out = IO.order_forward(sdm, rgb, np.stack(m1), np.stack(m2), np.zeros(2 * P, np.int64))
    print("oracle vs reference: depth %.2e occ %.2e" % (np.abs(out["depth"].reshape(P, 2, 3) - dl).max(),
                                                        np.abs(out["occ"].reshape(P, 2, 2) - ol).max()))


def disp_digest(disp):
    """Small fingerprint of a [384, 384] disparity map: 4x4 block means + three full rows."""
    d = np.asarray(disp, dtype=np.float64).reshape(D // 4, 4, D // 4, 4).mean(axis=(1, 3)).astype(np.float32)
    return d, np.asarray(disp, dtype=np.float32)[[0, D // 2 - 1, D - 1]]


def main_disp():
    """tests/golden/instadepth_disp.npz: the reference's disparity output on the fixture image, with the decoder /
    encoder-layer4 tensors of ``synth.instadepth_state_dict(with_decoder=True)`` loaded as well."""
    import torch
    import torchvision
    torch.hub.load = lambda *a, **k: torchvision.models.resnext101_32x8d(weights=None)
    ref_shim.load()
    midas_net = sys.modules["_instaorder_ref.midas.midas_net"]
    sd = IO.load_calibrated(os.path.join(GOLDEN, "instadepth_calib.npz"), SEED, with_decoder=True)
    net = midas_net.InstaDepthNet_od(path=None)
    missing, unexpected = net.load_state_dict({k[7:]: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not unexpected and all(("fc." in k) or ("num_batches_tracked" in k) for k in missing), missing[:5]
    net.eval()
    from oracle import oracle as O
    image, _, _ = build_scene()
    rgb = O.resize_mode_rgb(image, D)[None]
    zero = torch.zeros(1, 1, D, D)
    with torch.no_grad():
        disp = net(torch.from_numpy(rgb), zero, zero)[0][0].numpy()
    pooled, rows = disp_digest(disp)
    np.savez_compressed(os.path.join(GOLDEN, "instadepth_disp.npz"), pooled=pooled, rows=rows,
                        stats=np.asarray([disp.min(), disp.max(), disp.mean(), disp.std()], np.float64))
    mine = IO.disparity_forward(sd, rgb)[0]
    print("disparity: range [%.3f, %.3f], oracle vs reference %.2e" % (disp.min(), disp.max(), np.abs(mine - disp).max()))
    # depth order from the disparity map (inference.py:589-599 with :79-104), both selection methods, through the
    # reference's own driver; plus the per-instance statistics the decisions rest on (for the tie margin of the test)
    import types
    ns = ref_shim.load()
    image, masks, boxes = build_scene()
    model = types.SimpleNamespace(model=net)
    extra = {}
    for sel in ("median", "mean"):
        order, clipped = ns.inference.infer_order_sup_depth(model, image, masks, boxes, "all", "InstaDepthNet_od", "resize",
                                                            D, sel)
        extra["order_" + sel] = order.astype(np.int64)
        extra["clipped_" + sel] = disp_digest(clipped.numpy())[0]
        pd = 1.0 / (torch.from_numpy(disp) + 1e-6)
        st = []
        for m in masks:
            mm = torch.from_numpy(O.resize_mode_mask(m, D).astype(bool))
            v = pd[mm]
            c = torch.clip(v, torch.quantile(v, 0.05), torch.quantile(v, 0.95))
            st.append(float(torch.median(c) if sel == "median" else torch.mean(c)))
        extra["stat_" + sel] = np.asarray(st, np.float64)
        print(sel, order.tolist(), st)
    np.savez_compressed(os.path.join(GOLDEN, "instadepth_disp.npz"), pooled=pooled, rows=rows,
                        stats=np.asarray([disp.min(), disp.max(), disp.mean(), disp.std()], np.float64), **extra)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "disp":
        main_disp()
    else:
        main()
