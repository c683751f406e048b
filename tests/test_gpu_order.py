"""End-to-end parity through the reference-facing API: order matrices identical to the *reference's* on every
pair whose decision margin exceeds 1e-3, logits within 2e-2 absolute of the reference's fp32 logits
(BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from instaorder_b200 import engine, inference, models
from oracle import calib, gen_golden

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2       # |logit - reference fp32 logit|, BASELINE.json north_star
BF16_EMU_TOL = 5e-3    # |logit - ideal-bf16 CPU emulation of the same arithmetic| (kernel correctness proper)


def make_model(case, golden_dir, max_pairs=32):
    c = gen_golden.CASES[case]
    params = dict(algo=c["algo"], backbone_arch="resnet50_cls",
                  backbone_param=dict(in_channels=5, num_classes=c["num_classes"]), optim="SGD", lr=1e-4,
                  weight_decay=1e-4, use_rgb=True, max_pairs=max_pairs)
    m = models.__dict__[c["algo"]](params, dist_model=False)
    m.load_state_dict(gen_golden.state_dict_for(case))
    m.switch_to("eval")
    return m


# c2_od_full / c3_ordernet_full: BASELINE.json's own sizes (10 instances -> 45 pairs; 15 instances at 1242 x 375 ->
# 105 pairs), fixtures from the live reference like all the others
@pytest.mark.parametrize("case", ["c1_o", "c2_od", "c3_ordernet", "c2_d", "c3_ordernet_ext", "c2_od_resize",
                                  "c1_o_image", "c2_od_full", "c3_ordernet_full"])
def test_order_matrices_match_reference(golden_dir, case):
    c = gen_golden.CASES[case]
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    image, masks, boxes = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(boxes, 3.0)
    mode = c.get("patch_or_image", "patch")
    D = c.get("input_size", 256)
    model = make_model(case, golden_dir)
    # details (logits, margins) through the engine ...
    eng = model.engine_for(D)
    r = eng.infer_scenes([engine.Scene(image, masks, bexp)], c["algo"], "all", mode, return_details=True)[0]
    heads = engine.heads_for(c["algo"], c["num_classes"])
    off = 0
    for h, (_, k, what) in enumerate(heads):
        ref = z["logits%d" % h]
        got = r["logits"][:, :, off:off + k]
        err = np.abs(got - ref).max()
        print("%s head %d: max |logit - reference fp32| = %.5f (logit std %.3f, err / std = %.3f)" %
              (case, h, err, ref.std(), err / ref.std()))
        assert err < LOGIT_TOL, "%s head %d: max |logit - reference| = %.4f" % (case, h, err)
        off += k
    N = masks.shape[0]
    for what in ("occ", "depth"):
        if what not in r:
            continue
        mg = np.full((N, N), np.inf)
        for (i, j), m in zip(r["pairs"], r["margin_" + what]):
            mg[i, j] = mg[j, i] = m
        ok = mg > 1e-3
        assert ok.sum() >= 0.8 * N * (N - 1)
        assert np.array_equal(r[what][ok], z[what][ok]), (case, what)
    # ... and the drop-in functions return the same matrices with the reference's signature / types
    if c["algo"] == "InstaOrderNet_od":
        occ, depth = inference.infer_order_sup_occ_depth(model, image, masks, bexp, "all", c["algo"], mode, D, "")
        assert np.array_equal(occ, r["occ"]) and np.array_equal(depth, r["depth"])
        assert occ.dtype == np.int64 and occ.shape == (N, N)
    elif c["algo"] == "InstaOrderNet_d":
        depth, disp = inference.infer_order_sup_depth(model, image, masks, bexp, "all", c["algo"], mode, D, "")
        assert disp is None and np.array_equal(depth, r["depth"])
    else:
        occ = inference.infer_order_sup_occ(model, image, masks, bexp, "all", c["algo"], mode, D)
        assert np.array_equal(occ, r["occ"])


def test_orig_mode_matches_reference(golden_dir):
    """``patch_or_image='orig'`` (reference inference.py:401-408): the network input is the image at
    its own size rounded to multiples of 32 -- 333 x 500 -> 320 x 512, NOT square -- through io_image_resize_rgb_hw /
    io_pair_gather_resize_hw / io_net_forward_pairs_hw.  Fixture: the unmodified reference (oracle/gen_golden_orig.py)."""
    from oracle import gen_golden_orig
    case = gen_golden_orig.CASE
    c = gen_golden.CASES[case]
    z = np.load(os.path.join(golden_dir, "order_c2_od_orig.npz"))
    image, masks, boxes = gen_golden.build_scene(case)
    model = make_model(case, golden_dir)
    eng = model.engine_for_orig(image.shape[0], image.shape[1])
    r = eng.infer_scenes([engine.Scene(image, masks, boxes)], c["algo"], "all", "orig", return_details=True)[0]
    assert tuple(z["net_input_shape"][2:]) == (engine.closest_multiple_of(image.shape[0]),
                                               engine.closest_multiple_of(image.shape[1])) == (320, 512)
    off = 0
    for h, k in enumerate((2, 3)):
        ref = z["logits%d" % h]
        err = float(np.abs(r["logits"][:, :, off:off + k] - ref).max())
        print("orig head %d: max |logit - reference fp32| = %.5f (logit std %.3f)" % (h, err, ref.std()))
        assert err < LOGIT_TOL, err
        off += k
    N = masks.shape[0]
    for what in ("occ", "depth"):
        mg = np.full((N, N), np.inf)
        for (i, j), m in zip(r["pairs"], r["margin_" + what]):
            mg[i, j] = mg[j, i] = m
        ok = mg > 1e-3
        assert np.array_equal(r[what][ok], z[what][ok]), what
    occ, depth = inference.infer_order_sup_occ_depth(model, image, masks, boxes, "all", c["algo"], "orig", 384, "")
    assert np.array_equal(occ, r["occ"]) and np.array_equal(depth, r["depth"])
    # a second image of another size through the same engine (plans are rebuilt), then the first one again
    from instaorder_b200 import synth
    img2, m2, b2 = synth.make_scene(np.random.RandomState(3), 427, 640, 3)
    r2 = model.engine_for_orig(427, 640).infer_scenes([engine.Scene(img2, m2, b2)], c["algo"], "all", "orig")[0]
    assert r2["occ"].shape == (3, 3)
    again = eng.infer_scenes([engine.Scene(image, masks, boxes)], c["algo"], "all", "orig", return_details=True)[0]
    assert np.array_equal(again["logits"], r["logits"])
    # and the square modes still work on an engine that has served `orig` calls
    sq = eng.infer_scenes([engine.Scene(img2, m2, engine.expand_bbox(b2, 3.0))], c["algo"], "all", "resize")[0]
    assert sq["depth"].shape == (3, 3)


def test_realistic_logit_scale(golden_dir):
    """The same network with heads 5 x larger (logit std 1.4 - 1.6, |logit| up to 7: the O(1) scale of a trained head)
    against the live reference's fixture.  The bf16-vs-fp32 distance is a RELATIVE quantity -- it scales with the head
    -- so north_star's absolute 2e-2 (met at the calibrated scale, logit std ~ 0.3) is NOT met here: measured on B200
    max |error| = 0.05 - 0.065 = 3.4 - 4.5 % of the logit std (printed below, recorded in DESIGN.md section 3).  What
    that costs in decisions is measured, not tuned away: a probability moves by at most |logit error| / 4 per
    direction, so an entry can only flip when the reference's own decision margin is below the error; entries whose
    margin exceeds the measured logit error must equal the reference's, and the flips among the near-ties in between
    (margin in (1e-3, error]) are counted and bounded (measured: 2 of 90 occlusion entries)."""
    case = "c2_od_big"
    c = gen_golden.CASES[case]
    z = np.load(os.path.join(golden_dir, "order_%s.npz" % case))
    image, masks, boxes = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(boxes, 3.0)
    model = make_model(case, golden_dir, max_pairs=64)
    r = model.engine_for(256).infer_scenes([engine.Scene(image, masks, bexp)], c["algo"], "all", "patch",
                                           return_details=True)[0]
    off = 0
    errs = {}
    for h, (_, k, what) in enumerate(engine.heads_for(c["algo"], c["num_classes"])):
        ref = z["logits%d" % h]
        err = float(np.abs(r["logits"][:, :, off:off + k] - ref).max())
        errs[what] = err
        print("%s head %d (%s): logit std %.3f, max |logit| %.3f, max |logit - reference fp32| = %.4f, err / std = "
              "%.4f" % (case, h, what, ref.std(), np.abs(ref).max(), err, err / ref.std()))
        assert ref.std() > 1.0, "the case is meant to have O(1) logits"
        assert err / ref.std() < 0.06                    # the relative distance of the calibrated-scale cases (3 - 8 %)
        assert err < LOGIT_TOL * c["head_scale"]         # = the absolute budget scaled with the head
        off += k
    N = masks.shape[0]
    offdiag = ~np.eye(N, dtype=bool)
    for what in ("occ", "depth"):
        mg = np.full((N, N), np.inf)
        for (i, j), m in zip(r["pairs"], r["margin_" + what]):
            mg[i, j] = mg[j, i] = m
        # probability-space bound of the logit error: |d sigmoid| <= err / 4 (occlusion, also after averaging the two
        # directions); a softmax probability moves by <= err / 2, the gap between two of them by <= err (depth)
        thr = 1.2 * errs[what] / 4 if what == "occ" else errs[what]
        sure = (mg > thr) & offdiag                      # the logit error cannot reach across this margin
        near = (mg > 1e-3) & (mg <= thr) & offdiag
        flips_near = int((r[what] != z[what])[near].sum())
        flips_all = int((r[what] != z[what])[offdiag].sum())
        print("%s %s: %d of %d entries have margin > %.3f (must match), %d near-ties with margin in (1e-3, %.3f]: %d "
              "of them differ from the reference; %d entries differ in total (exact ties included)" %
              (case, what, int(sure.sum()), N * (N - 1), thr, int(near.sum()), thr, flips_near, flips_all))
        assert sure.sum() >= 0.5 * N * (N - 1)
        assert np.array_equal(r[what][sure], z[what][sure]), (case, what)
        assert flips_all <= 0.05 * N * (N - 1)


def test_benchmarked_batch_matches_oracle(golden_dir):
    """The configuration bench.py measures -- OrderEngine(max_pairs=256), default fused schedule, ONE full 256-pair
    batch (256-pair chunks, >= 1000-tile persistent launches, full phase-B chunk) of C2-shaped scenes -- compared pair
    by pair with the CPU oracle: fp32 restatement of the reference (<= 2e-2) and the ideal-bf16 emulation of the
    kernels' own arithmetic contract (<= 5e-3); order matrices equal on every pair off ties."""
    from instaorder_b200 import synth
    from oracle import oracle as O
    case = "c2_od"
    c = gen_golden.CASES[case]
    sd = gen_golden.state_dict_for(case)
    eng = engine.OrderEngine([2, 3], 256, max_pairs=256)
    eng.load_state_dict(sd)
    scenes, raw = [], []
    for k, (img, masks, boxes) in enumerate(synth.coco_scene_stream(41, 5, N=10)):       # 5 x 45 pairs
        raw.append((img, masks, boxes))
    rng = np.random.RandomState(43)
    raw.append(synth.make_scene(rng, 375, 500, 7, float_boxes=True))                      # + 21
    raw.append(synth.make_scene(rng, 333, 500, 5, float_boxes=True))                      # + 10 = 256
    for (img, masks, boxes) in raw:
        scenes.append(engine.Scene(img, masks, engine.expand_bbox(boxes, 3.0)))
    assert sum(s.n * (s.n - 1) // 2 for s in scenes) == 256
    launches0 = eng.gpu_launches
    res = eng.infer_scenes(scenes, c["algo"], "all", "patch", return_details=True)
    per_batch = eng.gpu_launches - launches0
    assert per_batch < 64, "expected ONE 256-pair batch, got %d launches" % per_batch
    worst32 = worst16 = 0.0
    nontie = total = 0
    for sc, r in zip(scenes, res):
        for fwd, tol in ((None, LOGIT_TOL), (O.resnet50_forward_bf16, BF16_EMU_TOL)):
            kw = {} if fwd is None else dict(forward=fwd)
            want = O.infer_order(sd, sc.image, sc.masks, sc.boxes.astype(np.int64), "all", c["algo"], "patch", 256, **kw)
            ref = np.stack([np.concatenate([np.stack(want["logits"][p]["fc_occ"]), np.stack(want["logits"][p]["fc_depth"])],
                                           axis=1) for p in want["pairs"]])
            assert [tuple(p) for p in want["pairs"]] == [tuple(p) for p in r["pairs"]]
            err = float(np.abs(r["logits"] - ref).max())
            assert err < tol, "max |logit - %s oracle| = %.4f" % ("fp32" if fwd is None else "ideal-bf16", err)
            if fwd is None:
                worst32 = max(worst32, err)
                for what in ("occ", "depth"):
                    ok = want["margin_" + what] > 1e-3
                    assert np.array_equal(r[what][ok], want[what][ok]), what
                    nontie += int(ok.sum()) - sc.n
                    total += sc.n * (sc.n - 1)
            else:
                worst16 = max(worst16, err)
    print("256-pair batch: max |logit - fp32 oracle| = %.5f, max |logit - ideal bf16| = %.5f, %d of %d matrix entries "
          "graded (non-tie), %d launches" % (worst32, worst16, nontie, total, per_batch))
    assert nontie >= 0.8 * total


@pytest.mark.parametrize("case", ["c2_od", "c3_ordernet"])
def test_logits_match_ideal_bf16_arithmetic(golden_dir, case):
    """The CUDA path against a CPU emulation of its own arithmetic contract (bf16 weights with BN folded, bf16
    stored activations, fp32 accumulation): only accumulation order differs, so the tolerance is much tighter
    than the bf16-vs-fp32 budget above."""
    from oracle import oracle as O
    c = gen_golden.CASES[case]
    image, masks, boxes = gen_golden.build_scene(case)
    bexp = engine.expand_bbox(boxes, 3.0)
    model = make_model(case, golden_dir)
    r = model.engine_for(256).infer_scenes([engine.Scene(image, masks, bexp)], c["algo"], "all", "patch",
                                           return_details=True)[0]
    sd = gen_golden.state_dict_for(case)
    want = O.infer_order(sd, image, masks, bexp, "all", c["algo"], "patch", 256, forward=O.resnet50_forward_bf16)
    names = ["fc_occ", "fc_depth"] if c["algo"] == "InstaOrderNet_od" else ["fc"]
    ref = np.stack([np.concatenate([np.stack(want["logits"][p][h]) for h in names], axis=1) for p in want["pairs"]])
    err = float(np.abs(r["logits"] - ref).max())
    print("%s: max |logit - ideal bf16| = %.5f" % (case, err))
    assert err < BF16_EMU_TOL, err


def test_nbor_pairs_and_multi_image_batching(golden_dir):
    """pairs='nbor' skips non-bordering pairs (their entries stay 0); batching several images / splitting one
    image over several batches gives the same matrices as one image at a time."""
    case = "c2_od"
    c = gen_golden.CASES[case]
    model = make_model(case, golden_dir)
    eng = model.engine_for(256)
    from instaorder_b200 import synth
    from oracle import oracle as O
    rng = np.random.RandomState(21)
    scenes = []
    for n in (2, 9, 3, 7):           # 1 + 36 + 3 + 21 pairs with max_pairs = 32 -> images split across batches
        img, masks, boxes = synth.make_scene(rng, 200, 260, n, wh_range=((20, 120), (20, 100)))
        scenes.append(engine.Scene(img, masks, engine.expand_bbox(boxes, 3.0)))
    together = eng.infer_scenes(scenes, c["algo"], "all", "patch")
    for sc, r in zip(scenes, together):
        alone = eng.infer_scenes([sc], c["algo"], "all", "patch")[0]
        assert np.array_equal(alone["occ"], r["occ"]) and np.array_equal(alone["depth"], r["depth"])
    sc = scenes[1]
    nb = eng.infer_scenes([sc], c["algo"], "nbor", "patch")[0]
    full = together[1]
    for i in range(sc.n):
        for j in range(i + 1, sc.n):
            if O.bordering(sc.masks[i], sc.masks[j]):
                assert nb["occ"][i, j] == full["occ"][i, j] and nb["depth"][i, j] == full["depth"][i, j]
                assert nb["occ"][j, i] == full["occ"][j, i] and nb["depth"][j, i] == full["depth"][j, i]
            else:
                assert nb["occ"][i, j] == 0 and nb["occ"][j, i] == 0 and nb["depth"][i, j] == 0
    # empty / single-instance images are legal and give empty / 1x1 zero matrices
    e = eng.infer_scenes([engine.Scene(scenes[0].image, scenes[0].masks[:1], scenes[0].boxes[:1])], c["algo"])[0]
    assert e["occ"].shape == (1, 1) and e["occ"][0, 0] == 0


def test_fused_and_unfused_schedules_agree():
    """The fused launches (dual-source layer-first blocks, conv3 -> next conv1 back-to-back GEMMs, the phase A -> B
    hand-over) against the one-launch-per-convolution schedule (INSTAORDER_FUSE=0, INSTAORDER_FUSE_DS=0) on the same
    pairs: same arithmetic up to the bf16 rounding of the identity tensor, which only the unfused schedule stores."""
    from instaorder_b200 import synth
    rng = np.random.RandomState(5)
    image, masks, boxes = synth.make_scene(rng, 240, 320, 6, wh_range=((30, 140), (30, 120)))
    bexp = engine.expand_bbox(boxes, 3.0)
    sd = calib.load_calibrated(gen_golden.calib_path("c2_od"), gen_golden.CASES["c2_od"]["wseed"], 5, [2, 3])
    out = {}
    for name, env in (("fused", {}), ("plain", {"INSTAORDER_FUSE": "0", "INSTAORDER_FUSE_DS": "0"})):
        old = {k: os.environ.get(k) for k in ("INSTAORDER_FUSE", "INSTAORDER_FUSE_DS")}
        os.environ.update(env)
        try:
            eng = engine.OrderEngine([2, 3], 256, max_pairs=16)     # the switches are read by io_net_create
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        eng.load_state_dict(sd)
        out[name] = eng.infer_scenes([engine.Scene(image, masks, bexp)], "InstaOrderNet_od", "all", "patch",
                                     return_details=True)[0]
        out[name + "_launches"] = eng.gpu_launches
    assert out["fused_launches"] < out["plain_launches"] - 10
    # two bf16 realisations of the same network (the identity of the four layer-first blocks is rounded to bf16 only
    # in the unfused schedule): each sits within BF16_EMU_TOL of the ideal emulation, so they sit within twice that
    err = np.abs(out["fused"]["logits"] - out["plain"]["logits"]).max()
    print("fused vs unfused: max |logit difference| = %.5f, launches %d vs %d" % (err, out["fused_launches"],
                                                                                  out["plain_launches"]))
    assert err < 2 * BF16_EMU_TOL, "fused vs unfused logits differ by %.4f" % err
    for what in ("occ", "depth"):
        ok = np.minimum(out["fused"]["margin_" + what], out["plain"]["margin_" + what]) > 1e-3
        ij = out["fused"]["pairs"][ok]
        for (i, j) in ij:
            assert out["fused"][what][i, j] == out["plain"][what][i, j]
            assert out["fused"][what][j, i] == out["plain"][what][j, i]


def test_full_size_permutation_property(golden_dir):
    """BASELINE-size check without an oracle run (the fp32 CPU reference needs minutes for 256 pairs): at one full
    256-pair batch (23 instances -> 253 pairs, plus a second image) relabelling the instances must permute the order
    matrices -- pair (i, j) becomes (pi(i), pi(j)), possibly with A and B exchanged, which swaps the two directions of
    the network (reference inference.py:144-161 averages them, so the decision is symmetric).  Exact off ties."""
    case = "c2_od"
    c = gen_golden.CASES[case]
    from instaorder_b200 import synth
    sd = calib.load_calibrated(gen_golden.calib_path(case), c["wseed"], 5, c["num_classes"])
    eng = engine.OrderEngine([2, 3], 256, max_pairs=256)
    eng.load_state_dict(sd)
    rng = np.random.RandomState(77)
    n = 23
    img, masks, boxes = synth.make_scene(rng, 427, 640, n, wh_range=((40, 300), (40, 300)))
    img2, masks2, boxes2 = synth.make_scene(rng, 375, 500, 4, wh_range=((40, 200), (40, 200)))
    bexp = engine.expand_bbox(boxes, 3.0)
    extra = engine.Scene(img2, masks2, engine.expand_bbox(boxes2, 3.0))
    base = eng.infer_scenes([engine.Scene(img, masks, bexp), extra], c["algo"], "all", "patch", return_details=True)[0]
    assert base["pairs"].shape[0] == 253
    perm = rng.permutation(n)                      # new index k holds old instance perm[k]
    r = eng.infer_scenes([extra, engine.Scene(img, masks[perm], bexp[perm])], c["algo"], "all", "patch",
                         return_details=True)[1]
    inv = np.argsort(perm)                         # old instance i sits at new index inv[i]
    checked = 0
    for what in ("occ", "depth"):
        mg = np.zeros((n, n))
        for (i, j), m in zip(base["pairs"], base["margin_" + what]):
            mg[i, j] = mg[j, i] = m
        for i in range(n):
            for j in range(n):
                if i != j and mg[i, j] > 1e-3:
                    assert r[what][inv[i], inv[j]] == base[what][i, j], (what, i, j)
                    checked += 1
    assert checked >= 0.8 * 2 * n * (n - 1)


def test_infer_stream_equals_blocking_calls(golden_dir):
    """``OrderEngine.infer_stream`` (two calls in flight, full-size first batches, results through pinned buffers) returns
    exactly what one blocking ``infer_scenes`` per call returns -- same kernels on the same inputs, so bit-identical."""
    case = "c2_od"
    c = gen_golden.CASES[case]
    from instaorder_b200 import synth
    sd = calib.load_calibrated(gen_golden.calib_path(case), c["wseed"], 5, c["num_classes"])
    eng = engine.OrderEngine([2, 3], 256, max_pairs=64)
    eng.load_state_dict(sd)
    rng = np.random.RandomState(5)
    calls = []
    for k in range(5):
        scenes = []
        for _ in range(1 + k % 3):
            img, masks, boxes = synth.make_scene(rng, 300 + 16 * k, 400, 3 + (k * 5) % 9, wh_range=((30, 200), (30, 200)))
            scenes.append(engine.Scene(img, masks, engine.expand_bbox(boxes, 3.0)))
        calls.append(scenes)
    want = [eng.infer_scenes(s, c["algo"]) for s in calls]
    got = list(eng.infer_stream(calls, c["algo"], depth=2))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert len(a) == len(b)
        for ra, rb in zip(a, b):
            for what in ("occ", "depth"):
                assert np.array_equal(ra[what], rb[what])
    # a handle collected late (other calls submitted in between) still holds its own result
    h0 = eng.submit_scenes(calls[0], c["algo"])
    h1 = eng.submit_scenes(calls[1], c["algo"])
    r1, r0 = eng.collect(h1), eng.collect(h0)
    for ra, rb in zip(r0, want[0]):
        assert np.array_equal(ra["occ"], rb["occ"]) and np.array_equal(ra["depth"], rb["depth"])
    for ra, rb in zip(r1, want[1]):
        assert np.array_equal(ra["occ"], rb["occ"]) and np.array_equal(ra["depth"], rb["depth"])
