"""Parity of the CUDA training step (io_train_* through instaorder_b200.training / models) with
  (a) a torch-autograd emulation of the same network with bf16 storage at the same points, teacher-forced layer by
      layer to the CUDA path's own forward state (kernel correctness: every stored tensor within 2 bf16 spacings,
      every gradient within 4e-2 relative L2 / cosine >= 0.999, loss 1e-3 relative),
  (b) the UNMODIFIED reference's ``step()`` frozen in tests/golden/train_*.npz (fp32): loss, logits, gradient norms,
      full gradients of five tensors, updated parameters, running statistics -- each within max(stated tolerance,
      2 x the deviation of an ideal bf16-storage emulation run freely from the same inputs),
  (c) the reference-shaped Python API: switch_to / set_input / step / forward_only / save_state / load_state(resume).
"""
import os
import tempfile

import numpy as np
import pytest
import torch

from instaorder_b200 import models, synth, training
from oracle import gen_golden_train as G
from oracle import train_oracle as T
import gpu_util as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _heads(algo):
    """(occ_off, class_off, class_k) for io_train_forward_backward."""
    return {"InstaOrderNet_od": (0, 2, 3), "InstaOrderNet_d": (-1, 0, 3), "InstaOrderNet_o": (0, -1, 0),
            "OrderNet": (-1, 0, 3)}[algo]


def _targets(algo, batch):
    occ = batch["occ_order"].to(DEV).float().contiguous() if algo in ("InstaOrderNet_od", "InstaOrderNet_o") else None
    if algo == "OrderNet":
        cls = batch["occ_order"].to(DEV).contiguous()
    elif algo in ("InstaOrderNet_od", "InstaOrderNet_d"):
        cls = batch["depth_order"].to(DEV).contiguous()
    else:
        cls = None
    ovl = batch["is_overlap"].to(DEV).contiguous() if T.ALGOS[algo][3] else None
    return occ, cls, ovl


def _run_engine(c, it_batches=1):
    algo = c["algo"]
    nc = T.ALGOS[algo][0]
    sd = synth.random_state_dict(c["wseed"], 5, nc)
    eng = training.TrainEngine(nc, c["D"], c["B"], DEV)
    eng.load_state_dict(sd)
    batch = T.make_batch(c["bseed"], c["B"], c["D"], algo)
    eng.pack_inputs(batch["rgb"], batch["modal1"], batch["modal2"])
    occ, cls, ovl = _targets(algo, batch)
    o, k, kk = _heads(algo)
    losses = eng.forward_backward(o, k, kk, occ, cls, ovl, c.get("overlap_weight", 1.0), c.get("distinct_weight", 1.0),
                                  1).clone()
    torch.cuda.synchronize()
    return eng, sd, batch, losses


def _emulate(c, sd, batch, eng=None):
    """bf16-storage torch emulation of the step; with ``eng`` every stored tensor is teacher-forced to the CUDA
    path's own value.  Returns (loss, logits [2,B,K], grads dict, running stats dict, per-tensor deviations)."""
    algo = c["algo"]
    nc = T.ALGOS[algo][0]
    names = T.param_names(nc)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    P = {k: torch.from_numpy(np.asarray(sd["module." + k])).to(DEV).requires_grad_(True) for k in names}
    S = {k[7:]: torch.from_numpy(np.asarray(v)).to(DEV).clone() for k, v in sd.items()
         if k.endswith(("running_mean", "running_var"))}
    bt = {k: v.to(DEV) for k, v in batch.items()}
    B = c["B"]
    xs = (torch.cat([bt["modal1"], bt["modal2"], bt["rgb"]], 1), torch.cat([bt["modal2"], bt["modal1"], bt["rgb"]], 1))
    dev_log = []
    outs = []
    for direction, x in enumerate(xs):
        forced = None
        if eng is not None:
            forced = {}
            for key in U.stored_keys():
                name, which = key.rsplit(".", 1)
                t = eng.activation(name, 0 if which == "y" else 1).float()
                t = t.view(2, B, -1)[direction]
                forced[key] = t          # reshaped lazily below (needs the channel count)
            shapes = {}
            with torch.no_grad():
                U.emulated_forward_train({k: v.detach() for k, v in P.items()},
                                         {k: v.clone() for k, v in S.items()}, x, record=shapes)
            for key in forced:
                n, ch, h, w = shapes[key].shape
                forced[key] = forced[key].view(n, h, w, ch).permute(0, 3, 1, 2).contiguous()
        outs.append(U.emulated_forward_train(P, S, x, forced=forced, dev_log=dev_log))
    loss, occ_loss, cls_loss = T.step_loss(algo, outs[0], outs[1], bt, c.get("overlap_weight", 1.0),
                                           c.get("distinct_weight", 1.0), 1)
    loss.backward()
    heads = [h for h in ("fc", "fc_occ", "fc_depth") if h in outs[0]]
    logits = torch.stack([torch.cat([o[h] for h in heads], 1) for o in outs]).detach()
    return float(loss.detach()), logits, {k: P[k].grad for k in names}, S, dev_log


CASES = ["od_sgd", "d_sgd", "o_sgd", "ordernet_sgd", "od_sgd_128"]


@pytest.mark.parametrize("case", CASES + ["od_sgd_384"])
def test_step_vs_bf16_emulation(case):
    """Kernel correctness, layer by layer.  Forward: every stored tensor equals what torch computes from the CUDA
    path's own inputs of that layer within 2.5 spacings of 2^-8 * max(|value|, 1), i.e. one bf16 ulp for values in
    [1, 2) (conv: fp32 accumulation order; BN: statistics in a different order).  Backward: all 163 gradients within 4e-2 relative L2 (and cosine >= 0.999) of autograd
    evaluated at the same forward state with bf16-rounded stored gradients.  Measured (tools/train_debug_bwd.py): the
    error grows smoothly from 0 at the FC heads to 1.4 % at conv1.weight / 3.4 % at the small-norm bn1.bias -- bf16
    rounding of ~100 stored gradient tensors in sequence, no step at any layer type."""
    c = G.CASES[case]
    nc = T.ALGOS[c["algo"]][0]
    eng, sd, batch, losses = _run_engine(c)
    loss, logits, grads_ref, S, dev_log = _emulate(c, sd, batch, eng)
    worst_fwd = max(dev_log, key=lambda kv: kv[1])
    assert worst_fwd[1] <= 2.5, "forward tensor %s deviates by %.2f bf16 spacings" % worst_fwd
    got_logits = eng.logits()
    assert float((got_logits - logits).abs().max()) < 2e-3, "logits differ by %.4g" % float(
        (got_logits - logits).abs().max())
    assert abs(float(losses[0]) - loss) <= 1e-3 * abs(loss), (float(losses[0]), loss)
    grads = eng.export_flat(eng.grads, params_only=True)
    worst = (0.0, None)
    for k in T.param_names(nc):
        g, w = grads[k].to(DEV), grads_ref[k]
        rel = float((g - w).norm() / (w.norm() + 1e-12))
        cos = float((g * w).sum() / (g.norm() * w.norm() + 1e-20))
        assert cos >= 0.999, "gradient of %s: cosine %.5f" % (k, cos)
        if rel > worst[0]:
            worst = (rel, k)
    print("%s: worst forward deviation %.2f spacings (%s); worst gradient rel-L2 error %.4g (%s)" % (
        case, worst_fwd[1], worst_fwd[0], worst[0], worst[1]))
    assert worst[0] <= 4e-2, "gradient of %s: relative L2 error %.4g" % (worst[1], worst[0])
    st = eng.export_flat(None, stats=eng.stats)
    for k, v in S.items():
        rel = float((st[k].to(DEV) - v).abs().max() / (v.abs().max() + 1e-6))
        assert rel < 2e-3, "running stat %s differs by %.4g" % (k, rel)


@pytest.mark.parametrize("case", CASES)
def test_step_vs_reference_golden(case, golden_dir):
    """Against the UNMODIFIED reference's fp32 step.  This synthetic random-init net with 4-sample BatchNorm batches
    amplifies bf16 rounding chaotically (x1.5 per bottleneck, tools/train_debug.py), so the bar is set by an *ideal*
    bf16-storage emulation run freely from the same inputs: the CUDA path may deviate from the reference by at most
    max(north_star's 2e-2, 2 x the ideal emulation's own deviation)."""
    c = G.CASES[case]
    algo = c["algo"]
    nc = T.ALGOS[algo][0]
    z = np.load(os.path.join(golden_dir, "train_%s.npz" % case))
    eng, sd, batch, losses = _run_engine(c)
    e_loss, e_logits, e_grads, _, _ = _emulate(c, sd, batch, None)
    want = float(z["s0_loss"])
    tol = max(2e-2 * abs(want), 2 * abs(e_loss - want))
    assert abs(float(losses[0]) - want) <= tol, "loss %.5f vs reference %.5f" % (float(losses[0]), want)
    lg = eng.logits().cpu().numpy()
    err = float(np.abs(lg - z["s0_logits"]).max())
    ideal = float(np.abs(e_logits.cpu().numpy() - z["s0_logits"]).max())
    print("%s: max |logit - reference| = %.4f (ideal bf16 emulation: %.4f)" % (case, err, ideal))
    assert err <= max(2e-2, 2 * ideal), "logits differ from the reference by %.4g (ideal bf16: %.4g)" % (err, ideal)
    names = T.param_names(nc)
    grads = eng.export_flat(eng.grads, params_only=True)
    # gradient norms: the vector of all per-tensor norms within 10 % (relative L2); single tensors within 25 % or
    # 3 x the ideal emulation's own deviation (individual BN gradients of this chaotic net move by 10-25 % under ANY
    # bf16 realisation -- the strict per-gradient check is test_step_vs_bf16_emulation)
    gn = np.array([float(grads[k].double().norm()) for k in names])
    en = np.array([float(e_grads[k].double().norm()) for k in names])
    wn = z["s0_grad_norm"]
    rel_vec = float(np.linalg.norm(gn - wn) / np.linalg.norm(wn))
    assert rel_vec <= 0.10, "gradient-norm vector deviates by %.3f from the reference" % rel_vec
    bad = np.abs(gn - wn) > np.maximum(0.25 * wn, 3 * np.abs(en - wn)) + 1e-7
    assert not bad.any(), "|grad %s| = %.5g vs reference %.5g (ideal bf16 %.5g)" % (
        names[int(np.argmax(bad))], gn[np.argmax(bad)], wn[np.argmax(bad)], en[np.argmax(bad)])
    for k in G.FULL_GRADS:
        w = torch.from_numpy(z["s0_fullgrad_" + k])
        rel = float((grads[k] - w).norm() / (w.norm() + 1e-12))
        rel_e = float((e_grads[k].cpu() - w).norm() / (w.norm() + 1e-12))
        assert rel <= max(8e-2, 2 * rel_e), "gradient %s: rel L2 error %.4g vs the reference (ideal bf16 %.4g)" % (
            k, rel, rel_e)
    # optimiser step, then compare the updated parameters and running statistics with the reference's
    opt = training.FlatOptim("SGD", G.LR, weight_decay=G.WEIGHT_DECAY)
    opt.attach(eng)
    opt.step()
    torch.cuda.synchronize()
    new = eng.state_dict()
    for i, k in enumerate(names):
        ps = T.tensor_digest(new["module." + k].numpy())[2]
        gscale = float(z["s0_grad_norm"][i]) / np.sqrt(new["module." + k].numel())
        tol = G.LR * (5 * gscale + 1e-6) + 1e-6
        assert float(np.abs(ps - z["s0_param_samples"][i]).max()) <= tol, "updated %s" % k
    rs = [k for k in new if k.endswith(("running_mean", "running_var"))]
    got = np.array([T.tensor_digest(new[k].numpy())[0] for k in rs])
    assert np.all(np.abs(got - z["s0_stat_norm"]) <= 3e-2 * z["s0_stat_norm"] + 1e-6)
    assert int(new["module.bn1.num_batches_tracked"]) == int(z["s0_nbt"])


def test_model_api_train_eval_resume(golden_dir):
    """models.InstaOrderNet_od: switch_to('train') / set_input / step x2 / switch_to('eval') / forward_only /
    save_state / load_state(resume=True), against the reference's two-step SGD fixture."""
    c = G.CASES["od_sgd"]
    z = np.load(os.path.join(golden_dir, "train_od_sgd.npz"))
    params = dict(G.case_params(c), device=DEV)
    m = models.InstaOrderNet_od(params)
    m.load_state_dict(synth.random_state_dict(c["wseed"], 5, [2, 3]))
    m.switch_to("train")
    losses = []
    for it in range(2):
        b = T.make_batch(c["bseed"] + it, c["B"], c["D"], c["algo"])
        m.set_input(**G.set_input_args(c["algo"], b))
        log, out = m.step()
        losses.append(float(out["loss"]))
        assert set(log) == {"loss_occ", "loss_depth"}
    assert abs(losses[0] - float(z["s0_loss"])) <= 2e-2 * float(z["s0_loss"])
    # second step: after an lr = 1e-2 update of this chaotic net only the magnitude is comparable (15 %)
    assert abs(losses[1] - float(z["s1_loss"])) <= 0.15 * float(z["s1_loss"]), (losses[1], float(z["s1_loss"]))
    # eval mode uses the updated weights (folded running statistics)
    m.switch_to("eval")
    log, out = m.forward_only()
    assert np.isfinite(float(out["loss"]))
    with tempfile.TemporaryDirectory() as td:
        m.save_state(td, 2)
        ck = torch.load(os.path.join(td, "ckpt_iter_2.pth.tar"), map_location="cpu", weights_only=False)
        assert set(ck) == {"step", "state_dict", "optimizer"} and ck["step"] == 2
        assert all(k.startswith("module.") for k in ck["state_dict"])
        assert len(ck["optimizer"]["state"]) == len(T.param_names([2, 3]))
        assert ck["optimizer"]["state"][0]["momentum_buffer"].shape == (64, 5, 7, 7)
        assert int(ck["state_dict"]["module.bn1.num_batches_tracked"]) == int(z["s1_nbt"])
        m2 = models.InstaOrderNet_od(params)
        m2.load_state(td, Iter=2, resume=True)
        m2.switch_to("train")
        m.switch_to("train")
        b = T.make_batch(c["bseed"] + 2, c["B"], c["D"], c["algo"])
        for mm in (m, m2):
            mm.set_input(**G.set_input_args(c["algo"], b))
        l1 = float(m.step()[1]["loss"])
        l2 = float(m2.step()[1]["loss"])
        assert abs(l1 - l2) <= 1e-3 * abs(l1), "resumed model diverges: %.6f vs %.6f" % (l1, l2)
        s1, s2 = m._trainer.state_dict(), m2._trainer.state_dict()
        for k in s1:
            if s1[k].dtype == torch.float32:
                d = float((s1[k] - s2[k]).abs().max())
                assert d <= 1e-3 * (float(s1[k].abs().max()) + 1e-3), "resume: %s differs by %.4g" % (k, d)


def test_trainer_loop(tmp_path):
    """instaorder_b200.trainer.Trainer (reference trainer.py:143-266): iteration loop with the LR schedule, loss
    recording, checkpoint cadence and on-line validation, on the synthetic dataset."""
    import types
    from instaorder_b200 import trainer as TR
    args = types.SimpleNamespace(
        seed=0, validate=False, load_pretrain=None, load_model=None,
        model=dict(G.case_params(G.CASES["od_sgd"]), device=DEV, total_iter=6, lr_steps=[4], lr_mults=[0.1], lr=1e-3,
                   warmup_lr=[], warmup_steps=[]),
        data=dict(dataset="synthetic", base_dir=str(tmp_path), patch_or_image="patch", batch_size=4, batch_size_val=4,
                  workers=0),
        trainer=dict(initial_val=False, val_freq=3, val_iter=2, print_freq=1, save_freq=3, loss_record=["loss"],
                     exp_name="unit"))
    tr = TR.Trainer(args, train_dataset=TR.SyntheticPairDataset("InstaOrderNet_od", 64, 64, seed=1),
                    val_dataset=TR.SyntheticPairDataset("InstaOrderNet_od", 64, 16, seed=2))
    tr.model.load_state_dict(synth.random_state_dict(20, 5, [2, 3]))
    tr.run()
    assert tr.curr_step == 6
    train_pts = [h for h in tr.history if "loss" in h[1]]
    val_pts = [h for h in tr.history if "val_loss" in h[1]]
    assert len(train_pts) == 6 and len(val_pts) == 2
    assert all(np.isfinite(h[1]["loss"]) for h in train_pts) and all(np.isfinite(h[1]["val_loss"]) for h in val_pts)
    assert abs(tr.model.optim.param_groups[0]["lr"] - 1e-4) < 1e-12          # milestone at iteration 4: lr x 0.1
    ck = os.path.join(tr.folder2save, "checkpoints")
    assert sorted(os.listdir(ck)) == ["ckpt_iter_3.pth.tar", "ckpt_iter_6.pth.tar"]
    st = torch.load(os.path.join(ck, "ckpt_iter_6.pth.tar"), map_location="cpu", weights_only=False)
    assert st["step"] == 6 and int(st["state_dict"]["module.bn1.num_batches_tracked"]) == 1 + 2 * 6


@pytest.mark.parametrize("name,algo", [("od", "InstaOrderNet_od"), ("o", "InstaOrderNet_o"), ("ordernet", "OrderNet")])
def test_train_batch_builder_matches_reference_getitem(name, algo, golden_dir):
    """G13 on the GPU: one fused gather launch builds the whole augmented batch (crop jitter / rescale, bicubic +
    nearest resize, flip, A/B swap, normalisation); its pair tensor equals bf16(reference ``__getitem__`` tensors)
    bit for bit, and the labels are the reference's."""
    from instaorder_b200 import engine, train_data as TD
    from oracle import gen_golden_traindata as GG
    z = np.load(os.path.join(golden_dir, "traindata.npz"))
    image, masks, boxes, occ, depth, overlap, count, geo = GG.make_scene()
    gt = dict(occ=occ, depth=depth, overlap=overlap, count=count)
    scene = engine.Scene(image, masks, boxes)
    B = GG.N_SAMPLES
    specs = []
    for k in range(B):
        np.random.seed(1000 + k)
        pair = None
        if name in ("od", "d"):
            s = geo[k % len(geo)]
            pair = tuple(map(int, s.split("<" if "<" in s else "=")))
        specs.append(TD.sample_pair(algo, boxes, gt, GG.BASE_AUG, pair=pair))
    bld = TD.TrainBatchBuilder(algo, GG.SZ, B, DEV)
    pt, labels = bld.build([scene] * B, specs)
    torch.cuda.synchronize()
    got, _ = U.unpack_pair_tensor(pt, B, GG.SZ)
    for k in range(B):
        want = U.f32_to_bf16_rn(z["%s_%d_x" % (name, k)])
        assert np.array_equal(got[k], want), "sample %d of %s differs (%d elements)" % (
            k, name, int((got[k] != want).sum()))
        assert np.array_equal(labels[k], z["%s_%d_labels" % (name, k)])
    # and the batch trains: set_input_pairs + step through the model API
    if name == "od":
        m = models.InstaOrderNet_od(dict(G.case_params(G.CASES["od_sgd"]), device=DEV))
        m.load_state_dict(synth.random_state_dict(20, 5, [2, 3]))
        m.switch_to("train")
        lab = torch.from_numpy(labels)
        keep = lab[:, 0] >= 0        # the reference's CE would reject label -1 as well; our synthetic GT has none
        assert bool(keep.all())
        m.set_input_pairs(pt, GG.SZ, lab[:, 0].long(), lab[:, 1].long(), lab[:, 2].long(), lab[:, 3:5].float())
        log, out = m.step()
        assert np.isfinite(float(out["loss"]))


@pytest.mark.parametrize("mode", ["resize", "image"])
def test_train_batch_builder_whole_image_modes(mode, golden_dir):
    """G13, `resize` (shipped ^od training config) and `image` modes on the GPU: per-scene u8 INTER_LINEAR plane (once) +
    one gather launch per batch (nearest masks, flip, A/B swap); masks bit-exact, rgb equal to bf16(reference)."""
    from instaorder_b200 import engine, train_data as TD
    from oracle import gen_golden_traindata as GG
    z = np.load(os.path.join(golden_dir, "traindata.npz"))
    image, masks, boxes, occ, depth, overlap, count, geo = GG.make_scene()
    gt = dict(occ=occ, depth=depth, overlap=overlap, count=count)
    scene = engine.Scene(image, masks, boxes)
    B = GG.N_SAMPLES
    specs = []
    for k in range(B):
        np.random.seed(2000 + k)
        s = geo[k % len(geo)]
        specs.append(TD.sample_pair("InstaOrderNet_od", boxes, gt, GG.BASE_AUG,
                                    pair=tuple(map(int, s.split("<" if "<" in s else "="))), mode=mode))
    bld = TD.TrainBatchBuilder("InstaOrderNet_od", GG.SZ, B, DEV, mode=mode)
    pt, labels = bld.build([scene] * B, specs)
    torch.cuda.synchronize()
    got, _ = U.unpack_pair_tensor(pt, B, GG.SZ)
    for k in range(B):
        want = U.f32_to_bf16_rn(z["od_%s_%d_x" % (mode, k)])
        assert np.array_equal(got[k], want), "%s sample %d differs (%d elements)" % (mode, k, int((got[k] != want).sum()))
        assert np.array_equal(labels[k], z["od_%s_%d_labels" % (mode, k)])
