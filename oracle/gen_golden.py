"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the *unmodified reference* into tests/golden/*.npz.

Run in the build container (needs /root/reference):  ``python -m oracle.gen_golden``.
The GPU box has no /root/reference, so the ``-m gpu`` parity tests compare the CUDA path with these fixtures
(and with oracle/oracle.py, which the ``not gpu`` tests pin against the same fixtures).

Fixtures (all inputs are re-generated from seeds by ``instaorder_b200.synth``; only reference *outputs* are stored):
  calib_<algo>.npz      calibrated BN statistics + FC heads of the synthetic checkpoint (oracle/calib.py)
  order_<case>.npz      reference order matrices, per-pair logits, crop boxes, gather-output digests + 3 full crops
  metrics.npz           reference eval_order_recall_precision_f1 / eval_depth_order_whdr on random matrices
  geometry.npz          reference combine_bbox / crop geometry / expand_bbox / crop_padding+cv2 on edge cases
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

from instaorder_b200 import synth
from oracle import calib, ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (algo, num_classes, weight seed, scene builder)
    "c1_o": dict(algo="InstaOrderNet_o", num_classes=2, wseed=1, scene=dict(seed=0, H=480, W=640, N=8),
                 expand=True, float_boxes=False),
    "c2_od": dict(algo="InstaOrderNet_od", num_classes=[2, 3], wseed=0, scene=dict(seed=5, H=427, W=640, N=6),
                  expand=True, float_boxes=True),
    "c3_ordernet": dict(algo="OrderNet", num_classes=3, wseed=2, scene=dict(seed=9, H=375, W=1242, N=5,
                        wh_range=((20, 200), (20, 150))), expand=True, float_boxes=False),
    "c2_d": dict(algo="InstaOrderNet_d", num_classes=3, wseed=3, scene=dict(seed=11, H=375, W=500, N=5),
                 expand=True, float_boxes=True),
    "c2_od_resize": dict(algo="InstaOrderNet_od", num_classes=[2, 3], wseed=5, scene=dict(seed=6, H=333, W=500, N=4),
                         expand=True, float_boxes=True, patch_or_image="resize", input_size=384),
    "c1_o_image": dict(algo="InstaOrderNet_o", num_classes=2, wseed=6, scene=dict(seed=17, H=375, W=500, N=5),
                       expand=True, float_boxes=False, patch_or_image="image", input_size=256),
    "c3_ordernet_ext": dict(algo="OrderNet", num_classes=4, wseed=4, scene=dict(seed=13, H=375, W=1242, N=4,
                            wh_range=((20, 200), (20, 150))), expand=True, float_boxes=False),
    # BASELINE.json's own sizes (round 2): C2 = 10 instances -> 45 pairs, C3 = 15 instances at 1242 x 375 -> 105 pairs
    "c2_od_full": dict(algo="InstaOrderNet_od", num_classes=[2, 3], wseed=0, scene=dict(seed=25, H=480, W=640, N=10),
                       expand=True, float_boxes=True),
    "c3_ordernet_full": dict(algo="OrderNet", num_classes=3, wseed=2, scene=dict(seed=29, H=375, W=1242, N=15,
                             wh_range=((20, 200), (20, 150))), expand=True, float_boxes=False),
    # realistic logit scale: the same checkpoint with both heads multiplied by 5 (logit std ~ 1.3 - 1.8 instead of
    # ~ 0.3); the bf16-vs-fp32 distance grows with it -- reported as |err| / std, not tuned to the absolute budget
    "c2_od_big": dict(algo="InstaOrderNet_od", num_classes=[2, 3], wseed=0, scene=dict(seed=25, H=480, W=640, N=10),
                      expand=True, float_boxes=True, head_scale=5.0),
}


def calib_path(case):
    c = CASES[case]
    nc = c["num_classes"]
    tag = "x".join(str(v) for v in nc) if isinstance(nc, list) else str(nc)
    return os.path.join(GOLDEN, "calib_w%d_nc%s.npz" % (c["wseed"], tag))


def state_dict_for(case):
    """The calibrated synthetic checkpoint of a case (numpy, reference key names): seeded conv weights + the frozen
    calibration tensors, heads multiplied by ``head_scale`` (one IEEE fp32 multiply: identical on every box)."""
    c = CASES[case]
    sd = calib.load_calibrated(calib_path(case), c["wseed"], 5, c["num_classes"])
    k = np.float32(c.get("head_scale", 1.0))
    if k != 1.0:
        for name in list(sd):
            if name.split(".")[-2] in ("fc", "fc_occ", "fc_depth"):
                sd[name] = (np.asarray(sd[name], dtype=np.float32) * k).astype(np.float32)
    return sd


def build_scene(case):
    c = CASES[case]
    s = dict(c["scene"])
    rng = np.random.RandomState(s.pop("seed"))
    image, masks, boxes = synth.make_scene(rng, float_boxes=c["float_boxes"], **s)
    return image, masks, boxes


def digest(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def make_reference_model(ns, algo, num_classes, sd_np):
    import torch
    params = dict(algo=algo, backbone_arch="resnet50_cls",
                  backbone_param=dict(in_channels=5, num_classes=num_classes),
                  optim="SGD", lr=1e-4, weight_decay=1e-4, use_rgb=True, overlap_weight=1.0, distinct_weight=1.0)
    model = ns.models.__dict__[algo](params, dist_model=False)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ckpt_iter_0.pth.tar")
        torch.save({"step": 0, "state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()},
                    "optimizer": model.optim.state_dict()}, path)
        model.load_state(td, Iter=0)      # the reference's own loader (models/single_stage_model.py:54-61)
    model.switch_to("eval")
    return model


def gen_calib():
    done = set()
    for case, c in CASES.items():
        p = calib_path(case)
        if p in done:
            continue
        done.add(p)
        sd = synth.random_state_dict(c["wseed"], 5, c["num_classes"])
        changed = calib.calibrate(sd, D=c.get("input_size", 256), mode=c.get("patch_or_image", "patch"))
        np.savez_compressed(p, **changed)
        print("wrote", p, sum(v.size for v in changed.values()), "floats")


def gen_order(ns, only=None):
    import torch
    infer = ns.inference
    for case, c in CASES.items():
        if only and case not in only:
            continue
        image, masks, boxes = build_scene(case)
        bexp = ns_expand(ns, boxes) if c["expand"] else boxes
        sd = state_dict_for(case)
        model = make_reference_model(ns, c["algo"], c["num_classes"], sd)
        mode = c.get("patch_or_image", "patch")
        D = c.get("input_size", 256)

        rec = dict(inputs=[], logits=[])
        def record(m, inp, out):
            rec["inputs"].append(inp[0].numpy().copy())
            rec["logits"].append([o.detach().numpy().copy() for o in (out if isinstance(out, tuple) else (out,))])

        hook = model.model.register_forward_hook(record)
        if c["algo"] == "InstaOrderNet_od":
            occ, depth = infer.infer_order_sup_occ_depth(model, image, masks, bexp, "all", c["algo"], mode, D, "")
        elif c["algo"] == "InstaOrderNet_d":
            depth, _ = infer.infer_order_sup_depth(model, image, masks, bexp, "all", c["algo"], mode, D, "")
            occ = np.zeros_like(depth)
        else:
            occ = infer.infer_order_sup_occ(model, image, masks, bexp, "all", c["algo"], mode, D)
            depth = np.zeros_like(occ)
        hook.remove()
        N = masks.shape[0]
        P = N * (N - 1) // 2
        assert len(rec["inputs"]) == 2 * P
        nheads = len(rec["logits"][0])
        logits = [np.stack([rec["logits"][k][h][0] for k in range(2 * P)]).reshape(P, 2, -1) for h in range(nheads)]
        x_first = np.stack([rec["inputs"][2 * k][0] for k in range(P)])        # [P,5,D,D] fp32, direction (A,B)
        out = dict(occ=occ.astype(np.int64), depth=depth.astype(np.int64), boxes_expanded=np.asarray(bexp))
        for h in range(nheads):
            out["logits%d" % h] = logits[h].astype(np.float32)
        # gather digests: masks exactly, rgb as the fp32 tensor the reference fed the network
        out["mask_digest"] = np.array([digest(x_first[k, :2].astype(np.uint8)) for k in range(P)], dtype=np.uint64)
        out["rgb_digest"] = np.array([digest(x_first[k, 2:]) for k in range(P)], dtype=np.uint64)
        keep = [0, P // 2, P - 1] if mode == "patch" else [P // 2]
        out["full_idx"] = np.array(keep)
        out["full_x"] = x_first[keep].astype(np.float32)
        # direction (B,A) must be the channel swap of (A,B): assert here once, for every pair
        for k in range(P):
            assert np.array_equal(rec["inputs"][2 * k + 1][0], rec["inputs"][2 * k][0][[1, 0, 2, 3, 4]])
        p = os.path.join(GOLDEN, "order_%s.npz" % case)
        np.savez_compressed(p, **out)
        print("wrote", p, "occ sum", int(occ.sum()), "depth hist", np.bincount(depth.reshape(-1), minlength=3))


def ns_expand(ns, boxes):
    """Tester.expand_bbox needs a Tester instance (tools/test.py:155-163); call the function body through a stub."""
    import importlib.util
    import types
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "tools", "test.py")).read()
    start = src.index("    def expand_bbox(self, bboxes):")
    end = src.index("    def run(self):")
    body = "import numpy as np\n" + "\n".join(l[4:] for l in src[start:end].splitlines())
    mod = types.ModuleType("_ref_expand")
    exec(compile(body, "tools/test.py[expand_bbox]", "exec"), mod.__dict__)
    self = types.SimpleNamespace(args=types.SimpleNamespace(data={"enlarge_box": 3.0}))
    return mod.expand_bbox(self, boxes)


def gen_metrics(ns):
    infer = ns.inference
    rng = np.random.RandomState(42)
    recs = dict(N=[], order=[], gt=[], zd=[], prf=[], depth_pred=[], gtd=[], ovl=[], cnt=[], whdr=[])
    for t in range(64):
        N = int(rng.randint(2, 17)) if t > 3 else [2, 2, 3, 3][t]
        occ_gt, depth_gt, ovl, cnt = synth.make_gt(rng, N)
        if t % 7 == 0:
            occ_gt[occ_gt == 1] = 0           # no positives in gt -> zero_division path
        if t % 5 == 0:
            ovl[:] = 0                        # empty ovlO masks -> -1
        if t % 11 == 0:
            depth_gt[depth_gt == 2] = 0
        pred = (rng.rand(N, N) < (0.0 if t % 9 == 0 else 0.3)).astype(np.int64)
        np.fill_diagonal(pred, 0)
        dpred = rng.randint(0, 3, size=(N, N)).astype(np.int64)
        zd = int(t % 2)
        prf = infer.eval_order_recall_precision_f1(pred, occ_gt, zd)
        wh = infer.eval_depth_order_whdr(dpred, (depth_gt, ovl, cnt))
        from oracle.oracle import WHDR_KEYS
        pad = lambda a: np.pad(a, ((0, 16 - N), (0, 16 - N)))
        recs["N"].append(N); recs["zd"].append(zd)
        recs["order"].append(pad(pred)); recs["gt"].append(pad(occ_gt)); recs["prf"].append(prf)
        recs["depth_pred"].append(pad(dpred)); recs["gtd"].append(pad(depth_gt)); recs["ovl"].append(pad(ovl))
        recs["cnt"].append(pad(cnt)); recs["whdr"].append([float(wh[k][0]) for k in WHDR_KEYS])
    p = os.path.join(GOLDEN, "metrics.npz")
    np.savez_compressed(p, **{k: np.array(v) for k, v in recs.items()})
    print("wrote", p)


def kins_scene(seed, n=7, H=96, W=160):
    """Modal / amodal masks: amodal = modal plus a dilated halo, so neighbouring instances overlap in amodal space."""
    rng = np.random.RandomState(seed)
    _, modal, _ = synth.make_scene(rng, H, W, n, wh_range=((15, 70), (15, 50)))
    # later instances are drawn on top: remove their pixels from earlier modal masks (a real occlusion pattern)
    for i in range(n):
        for j in range(i + 1, n):
            modal[i] &= ~modal[j] & 1
    amodal = modal.copy()
    for i in range(n):
        a = amodal[i]
        for _ in range(int(rng.randint(0, 6))):
            d = a.copy()
            d[1:] |= a[:-1]; d[:-1] |= a[1:]; d[:, 1:] |= a[:, :-1]; d[:, :-1] |= a[:, 1:]
            a = d
        amodal[i] = a
    return modal, amodal


def gen_gt_order(ns):
    recs = {}
    for t, seed in enumerate((3, 4, 5)):
        modal, amodal = kins_scene(seed)
        recs["gt%d" % t] = ns.inference.infer_gt_order(modal, amodal).astype(np.int64)
        recs["seed%d" % t] = seed
    p = os.path.join(GOLDEN, "gt_order.npz")
    np.savez_compressed(p, **recs)
    print("wrote", p, [int(recs["gt%d" % t].sum()) for t in range(3)])


def gen_geometry(ns):
    """Crop geometry + crop_padding + cv2 resizes on hand-picked and random edge cases, from the reference's own
    utils.combine_bbox / utils.crop_padding / inference.resize_mask and cv2 (IPP off, see ref_shim)."""
    import cv2
    utils, infer = ns.utils, ns.inference
    rng = np.random.RandomState(7)
    boxes, crops = [], []
    for t in range(200):
        if t < 100:
            b = np.array([[rng.randint(-50, 600), rng.randint(-50, 400), rng.randint(1, 400), rng.randint(1, 300)]
                          for _ in range(2)], dtype=np.int64)
        else:
            b = np.round(np.array([[rng.uniform(0, 600), rng.uniform(0, 400), rng.uniform(0.5, 400),
                                    rng.uniform(0.5, 300)] for _ in range(2)]), 2)
        bbox = utils.combine_bbox(b[(0, 1), :])
        cx = bbox[0] + bbox[2] / 2.
        cy = bbox[1] + bbox[3] / 2.
        size = max([np.sqrt(bbox[2] * bbox[3] * 2.), bbox[2] * 1.1, bbox[3] * 1.1])
        nb = [int(cx - size / 2.), int(cy - size / 2.), int(size), int(size)]
        boxes.append(b.astype(np.float64))
        crops.append(nb)
    image = rng.randint(0, 256, size=(120, 160, 3)).astype(np.uint8)
    mask = (rng.rand(120, 160) < 0.4).astype(np.uint8)
    rois = [[-20, -30, 90, 90], [100, 60, 200, 200], [10, 10, 37, 37], [-300, -300, 100, 100], [0, 0, 160, 160],
            [150, 110, 5, 5], [-5, 50, 1, 1], [40, 30, 256, 256], [40, 30, 255, 255], [20, 20, 2, 2],
            [-400, -200, 1000, 1000], [30, 30, 3, 3]]
    rgbs, ms = [], []
    for roi in rois:
        rgbs.append(cv2.resize(utils.crop_padding(image, roi, pad_value=(0, 0, 0)), (64, 64),
                               interpolation=cv2.INTER_CUBIC))
        ms.append(infer.resize_mask(utils.crop_padding(mask, roi, pad_value=(0,)), 64, "nearest"))
    p = os.path.join(GOLDEN, "geometry.npz")
    np.savez_compressed(p, boxes=np.array(boxes), crops=np.array(crops), image=image, mask=mask,
                        rois=np.array(rois), rgb=np.array(rgbs), m=np.array(ms))
    print("wrote", p)


def val_batch(seed, algo, num_classes, B=6, D=256):
    """C4-style synthetic collated batch: normal rgb, blob masks, random labels."""
    rng = np.random.RandomState(seed)
    rgb = rng.standard_normal((B, 3, D, D)).astype(np.float32)
    m = np.zeros((B, 2, D, D), dtype=np.float32)
    for b in range(B):
        for c in range(2):
            x0, y0 = rng.randint(0, D // 2, size=2)
            w, h = rng.randint(D // 8, D // 2, size=2)
            m[b, c, y0:y0 + h, x0:x0 + w] = 1.0
    out = dict(rgb=rgb, modal1=m[:, 0:1].copy(), modal2=m[:, 1:2].copy())
    if algo in ("InstaOrderNet_od", "InstaOrderNet_d"):
        out["depth_order"] = rng.randint(0, 3, size=B).astype(np.int64)
        out["count"] = rng.randint(2, 4, size=B).astype(np.int64)
        out["is_overlap"] = (rng.rand(B) < 0.4).astype(np.int64)
    if algo == "InstaOrderNet_od":
        out["occ_order"] = (rng.rand(B, 2) < 0.3).astype(np.float32)
    elif algo == "InstaOrderNet_o":
        out["occ_order"] = (rng.rand(B, 2) < 0.3).astype(np.float32)
    elif algo == "OrderNet":
        out["occ_order"] = rng.randint(0, int(num_classes), size=B).astype(np.int64)
    return out


VAL_CASES = {"c2_od": 31, "c1_o": 32, "c2_d": 33, "c3_ordernet": 34, "c3_ordernet_ext": 35}


def gen_losses(ns):
    """Reference set_input + forward_only (validation loss) and model.model(x) logits on synthetic batches."""
    import torch
    rec = {}
    for case, seed in VAL_CASES.items():
        c = CASES[case]
        sd = calib.load_calibrated(calib_path(case), c["wseed"], 5, c["num_classes"])
        model = make_reference_model(ns, c["algo"], c["num_classes"], sd)
        model.params["overlap_weight"], model.params["distinct_weight"] = 1.5, 0.5
        batch = val_batch(seed, c["algo"], c["num_classes"])
        model.set_input(**{k: torch.from_numpy(v) for k, v in batch.items()})
        r = model.forward_only()
        log, loss = r if isinstance(r, tuple) else ({}, r)
        rec[case + "_loss"] = np.float32(loss["loss"].item())
        for k, v in log.items():
            rec[case + "_" + k] = np.float32(v.item())
        with torch.no_grad():
            x = torch.cat([torch.from_numpy(batch["modal1"]), torch.from_numpy(batch["modal2"]),
                           torch.from_numpy(batch["rgb"])], dim=1)
            y = model.model(x)
            y2 = model.model(x[:, [1, 0, 2, 3, 4]])
        y = torch.cat(y, dim=1) if isinstance(y, tuple) else y
        y2 = torch.cat(y2, dim=1) if isinstance(y2, tuple) else y2
        rec[case + "_logits"] = torch.stack([y, y2], dim=1).numpy().astype(np.float32)     # [B, 2, K]
        print(case, {k: float(v) for k, v in rec.items() if k.startswith(case) and not k.endswith("logits")})
    p = os.path.join(GOLDEN, "val_losses.npz")
    np.savez_compressed(p, **rec)
    print("wrote", p)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ns = ref_shim.load()
    what = sys.argv[1:] or ["calib", "geometry", "metrics", "order", "losses", "gt"]
    if "calib" in what:
        gen_calib()
    if "geometry" in what:
        gen_geometry(ns)
    if "metrics" in what:
        gen_metrics(ns)
    if "order" in what:
        gen_order(ns, only=[w for w in what if w in CASES] or None)
    if "losses" in what:
        gen_losses(ns)
    if "gt" in what:
        gen_gt_order(ns)


if __name__ == "__main__":
    main()
