#!/usr/bin/env python
"""Benchmark of the pairwise-order hot path (BASELINE.json metric: instance pairs/s, InstaOrderNet^od, 256^2, bf16).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, one process per GPU (torchrun for N>1)
    python bench.py --impl reference ...                      # the reference algorithm on the box's host cores

One *step* = one pass of the hot path (fused gather -> 5-ch ResNet-50, both directions -> decide + scatter) over
one batch of 256 instance pairs cut from synthetic COCO-shaped images (10 instances -> 45 pairs / image, patch 256).
`value`  : whole-job pairs/s with the u8 images / masks / pair descriptors already resident in HBM.
`e2e`    : the same metric through the public API (`OrderEngine.infer_scenes`) from HOST numpy buffers: pinned staging,
           H2D copies and the D2H read of the order matrices are inside the timed region.
`roofline`: the tensor-core convolution kernel (conv_tc_kernel), algorithmic FLOPs / CUDA-event time per launch.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner on STDOUT (file descriptor 1, from C) at communicator creation whenever NCCL_DEBUG is
# VERSION or above: keep stdout to the ONE JSON line by pointing fd 1 at stderr for the whole run and writing the result
# line to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


FLOP_PER_PAIR = 21.764e9          # 2 x 10.882 GFLOP (conv + FC, 2*MAC) @256^2
PAIRS_PER_STEP = 256
ALGO = "InstaOrderNet_od"
NUM_CLASSES = [2, 3]
D = 256


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return dict(bf16=j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0)), hbm=j.get("hbm_gbs", 6650.0),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(bf16=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # clocks under load = upper half of the samples (the sampler also sees idle gaps)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return dict(sm_mhz=float(np.median(load)) if load else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_scenes(seed, n_images):
    from instaorder_b200 import engine, synth
    out = []
    for image, masks, boxes in synth.coco_scene_stream(seed, n_images, N=10):
        out.append(engine.Scene(image, masks, engine.expand_bbox(boxes, 3.0)))
    return out


def cpu_port_pairs_per_s(n_pairs, threads=None):
    """The oracle port (= the reference algorithm: per-pair cv2-equivalent crops, fp32 torch-CPU ResNet-50 on both
    directions, decisions) timed on the host cores for a bounded sample of the same workload."""
    import torch
    from instaorder_b200 import synth
    from oracle import oracle as O
    if threads is None:      # every host core this process may use (torchrun pins OMP_NUM_THREADS=1 by default)
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.RandomState(1234)
    sd = synth.random_state_dict(0, 5, NUM_CLASSES)
    # whole 10-instance images (45 pairs each) until n_pairs is reached; the last image is cut by dropping instances
    # (k instances -> k(k-1)/2 pairs)
    n_img = max(1, (n_pairs + 44) // 45)
    scenes = list(synth.coco_scene_stream(99, n_img, N=10))
    image, masks, boxes = scenes[0]
    bexp = O.expand_bbox(boxes, 3.0)
    O.infer_order(sd, image, masks[:2], bexp[:2], "all", ALGO, "patch", D)          # warm-up (1 pair)
    done, dt = 0, 0.0
    for (image, masks, boxes) in scenes:
        left = n_pairs - done
        if left <= 0:
            break
        k = 2
        while k * (k - 1) // 2 < left and k < 10:
            k += 1
        bexp = O.expand_bbox(boxes, 3.0)
        t0 = time.perf_counter()
        r = O.infer_order(sd, image, masks[:k], bexp[:k], "all", ALGO, "patch", D, chunk=8)
        dt += time.perf_counter() - t0
        done += len(r["pairs"])
    return done / dt, done, dt, torch.get_num_threads()


def cpu_port_instadepth(threads=None):
    """The oracle port of InstaDepthNet^od's order branch on the host cores: one 4-instance image (6 pairs)."""
    import torch
    from instaorder_b200 import synth
    from oracle import instadepth_oracle as IO, oracle as O
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.instadepth_state_dict(0)
    image, masks, _ = next(synth.coco_scene_stream(99, 1, N=4))
    rgb = O.resize_mode_rgb(image, 384)[None]
    mm = [O.resize_mode_mask(m, 384)[None].astype(np.float32) for m in masks]
    m1, m2 = [], []
    for (i, j) in O.enumerate_pairs(4):
        m1 += [mm[i], mm[j]]
        m2 += [mm[j], mm[i]]
    IO.order_forward(sd, rgb, np.stack(m1[:2]), np.stack(m2[:2]), np.zeros(2, np.int64))      # warm-up
    t0 = time.perf_counter()
    IO.order_forward(sd, rgb, np.stack(m1), np.stack(m2), np.zeros(len(m1), np.int64))
    dt = time.perf_counter() - t0
    return (len(m1) // 2) / dt, len(m1) // 2, dt, torch.get_num_threads()


def run_reference_arm(args):
    """`--impl reference`: the reference's algorithm on the host cores (the reference itself is Python and lives at
    /root/reference, which does not exist on the GPU box; oracle/oracle.py is its restatement, pinned against it by
    tests/test_oracle_golden.py).  Rank 0 only; other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps, warmup = args.steps, args.warmup
    sample_pairs = 10                                           # 5 instances -> 10 pairs per step
    vals = []
    for s in range(warmup + steps):
        v, n, dt, threads = cpu_port_pairs_per_s(sample_pairs)
        if s >= warmup:
            vals.append((n, dt))
    pairs = sum(n for n, _ in vals)
    secs = sum(dt for _, dt in vals)
    value = pairs / secs
    # same metric / config.workload strings as the b200 arm (the driver pairs the two lines); the reference arm computes in
    # fp32 (dtype) and its step is a bounded sample of the workload (config.sample)
    line = dict(impl="reference", metric="instance pairs/s (InstaOrderNet^od, 256^2, bf16)", value=value, unit="pairs/s",
                n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=1000.0 * secs / max(len(vals), 1),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="C2: synthetic COCO-val-shaped images, 10 instances -> 45 pairs/image, %d pairs per "
                                     "step, patch 256^2, InstaOrderNet^od heads [2,3], random-init weights" % PAIRS_PER_STEP,
                            sample="bounded sample of %d pairs per step on the host cores (fp32 torch-CPU oracle port of "
                                   "the reference's inference.py patch path)" % sample_pairs),
                cpu_baseline=dict(value=value, unit="pairs/s", cores=threads, kind="port",
                                  sample="%d pairs/step x %d steps, batched fp32 torch-CPU forward" % (sample_pairs, steps)),
                e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def cpu_train_pairs_per_s(n_pairs, d, threads=None):
    """The training oracle (= the reference's step(): two train-mode fp32 forwards, loss, backward, SGD) timed on
    the host cores for one small batch."""
    import torch
    from instaorder_b200 import synth
    from oracle import train_oracle as T
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.random_state_dict(0, 5, NUM_CLASSES)
    batch = T.make_batch(7, n_pairs, d, ALGO)
    t0 = time.perf_counter()
    T.train_step(sd, batch, ALGO, lr=1e-4, weight_decay=1e-4, overlap_weight=0.1, distinct_weight=0.9)
    dt = time.perf_counter() - t0
    return n_pairs / dt, n_pairs, dt, torch.get_num_threads()


def run_train(args):
    """BASELINE config C4: InstaOrderNet^od training step on synthetic pairs, per-GPU batch --train-batch, SGD
    (lr 1e-4, momentum 0.9, wd 1e-4: experiments/InstaOrder/InstaOrderNet_od/config.yaml), NCCL all-reduce of the
    flat gradient buffer for N > 1.  Not the headline metric: run explicitly with --workload train."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.train_batch
    if args.impl == "reference":
        if rank != 0:
            return
        vals = [cpu_train_pairs_per_s(2, D) for _ in range(args.warmup + args.steps)][args.warmup:]
        pairs, secs = sum(v[1] for v in vals), sum(v[2] for v in vals)
        value = pairs / secs
        emit((dict(impl="reference", metric="training pairs/s (InstaOrderNet^od step: fwd + bwd + all-reduce + SGD, 256^2, bf16)", value=value,
                              unit="pairs/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=1000.0 * secs / len(vals), higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype="f32", data="synthetic",
                              config=dict(workload="C4: InstaOrderNet^od step() on 2 synthetic pairs per step, host "
                                                   "cores (bounded sample)"),
                              cpu_baseline=dict(value=value, unit="pairs/s", cores=vals[0][3], kind="port",
                                                sample="2 pairs/step x %d steps, fp32 torch-CPU autograd" % len(vals)),
                              e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    import torch
    import torch.distributed as dist
    from instaorder_b200 import _lib, models, synth
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    params = dict(algo=ALGO, backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=NUM_CLASSES),
                  optim="SGD", lr=1e-4, weight_decay=1e-4, use_rgb=True, overlap_weight=0.1, distinct_weight=0.9,
                  device=dev)
    model = models.InstaOrderNet_od(params, dist_model=world > 1)
    model.load_state_dict(synth.random_state_dict(0, 5, NUM_CLASSES))
    model.switch_to("train")
    g = torch.Generator().manual_seed(100 + rank)
    n_batches = 4
    host = []
    for _ in range(n_batches):     # pinned host batches in the DataLoader's collated types
        host.append(dict(rgb=torch.randn((B, 3, D, D), generator=g).pin_memory(),
                         modal1=(torch.rand((B, 1, D, D), generator=g) > 0.7).float().pin_memory(),
                         modal2=(torch.rand((B, 1, D, D), generator=g) > 0.7).float().pin_memory(),
                         depth_order=torch.randint(0, 3, (B,), generator=g).pin_memory(),
                         count=torch.randint(2, 4, (B,), generator=g).pin_memory(),
                         is_overlap=(torch.rand((B,), generator=g) < 0.3).long().pin_memory(),
                         occ_order=(torch.rand((B, 2), generator=g) < 0.2).float().pin_memory()))
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(batch):
        model.set_input(**batch)
        return model.step()

    for i in range(args.warmup):
        step(resident[i % n_batches])
    sync_all()
    eng = model._trainer
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(resident[i % n_batches])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.gpu_launches - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * B / (ms / 1000.0)
    # end to end: pinned host batch -> H2D -> step -> loss read back, every step
    for i in range(2):
        float(step(host[i % n_batches])[1]["loss"])
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        float(step(host[i % n_batches])[1]["loss"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * args.steps * B / dt
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    # per-kernel events of one step
    _lib.check(eng.lib.io_train_profile(eng.handle, 1))
    step(resident[0])
    torch.cuda.synchronize()
    mx = 4096
    pms = np.zeros(mx, np.float32); kind = np.zeros(mx, np.int32); fl = np.zeros(mx, np.float64)
    by = np.zeros(mx, np.float64)
    n = _lib.check(eng.lib.io_train_profile_read(eng.handle, _lib.ptr(pms), _lib.ptr(kind), _lib.ptr(fl), _lib.ptr(by),
                                                 None, mx))
    _lib.check(eng.lib.io_train_profile(eng.handle, 0))
    peaks = measured_peaks()
    tc = kind[:n] <= 2
    ew = kind[:n] == 3
    tc_ms, tc_fl = float(pms[:n][tc].sum()), float(fl[:n][tc].sum())
    ew_ms, ew_by = float(pms[:n][ew].sum()), float(by[:n][ew].sum())
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, npairs, cdt, threads = cpu_train_pairs_per_s(2, D)
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind="port",
                       sample="one step() on 2 pairs (%.1f s): training oracle, fp32 torch-CPU autograd" % cdt)
        achieved = tc_fl / (tc_ms / 1000.0) / 1e12 if tc_ms > 0 else 0.0
        emit((dict(
            metric="training pairs/s (InstaOrderNet^od step: fwd + bwd + all-reduce + SGD, 256^2, bf16)",
            value=value, unit="pairs/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
            data="synthetic",
            config=dict(workload="C4: InstaOrderNet^od training step, %d synthetic pairs per GPU per step at %d^2, SGD "
                                 "lr 1e-4 momentum 0.9 wd 1e-4, random-init weights" % (B, D),
                        pairs_per_step=B * world, parallelism="data parallel over %d GPU(s), one NCCL all-reduce of the "
                        "flat fp32 gradient buffer (94 MB) per step" % world,
                        l2="each step streams > 10 GB of saved activations and gradients, i.e. >> 126 MB L2"),
            e2e=dict(value=e2e_value, unit="pairs/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4),
            gpu_launches=launches, clocks=clocks,
            roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                          frac=achieved / peaks["bf16"], traffic=None,
                          kernel="conv_tc / conv_tn (forward + data gradient) + wgrad_kernel",
                          tensor_share_of_step=tc_ms / float(pms[:n].sum()), peak_source=peaks["source"],
                          elementwise=dict(bound="hbm", achieved=ew_by / (ew_ms / 1000.0) / 1e9 if ew_ms else 0.0,
                                           peak=peaks["hbm"], unit="GB/s",
                                           frac=ew_by / (ew_ms / 1000.0) / 1e9 / peaks["hbm"] if ew_ms else 0.0,
                                           share_of_step=ew_ms / float(pms[:n].sum())),
                          step_frac=value / world * 3 * FLOP_PER_PAIR / 1e12 / peaks["bf16"]),
            cpu_baseline=cpu)))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-pairs", type=int, default=225)
    ap.add_argument("--mode", default="patch256", choices=["patch256", "resize384", "instadepth384"],
                    help="patch256 = the BASELINE.json metric (default); resize384 = the shipped InstaOrderNet^od "
                         "config (whole image -> 384^2), reported as a second row in DESIGN.md")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = the BASELINE.json headline (default); train = BASELINE config C4, one "
                         "InstaOrderNet^od training step (fwd + bwd + all-reduce + SGD) per step")
    ap.add_argument("--train-batch", type=int, default=32, help="pairs per GPU per training step (reference: 32)")
    args = ap.parse_args()
    if args.workload == "train":
        return run_train(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from instaorder_b200 import _lib, engine, synth
    global D, FLOP_PER_PAIR, PAIRS_PER_STEP, ALGO
    gmode = "patch"
    depth = args.mode == "instadepth384"      # BASELINE config 5: InstaDepthNet^od order inference, 384^2
    if args.mode == "resize384":
        D, FLOP_PER_PAIR, gmode = 384, 48.970e9, "resize"
    if depth:
        from instaorder_b200 import depth_engine
        D, gmode, PAIRS_PER_STEP, ALGO = 384, "resize", 128, "InstaDepthNet_od"
        # our formulation: two trunks x two directions per pair + the encoder once per image (45 pairs per image);
        # the reference spends 2 x 254.5 GFLOP per pair (encoder + MiDaS decoder + trunks for every direction)
        FLOP_PER_PAIR = 4 * 24.485e9 + depth_engine.encoder_flops_per_image(D) / 45.0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    # ---- workload: every rank owns its own images (weak scaling: images are sharded, no data-path collective)
    n_batches = 8                                    # distinct resident batches, rotated so inputs never sit in L2
    n_images = (n_batches * PAIRS_PER_STEP) // 45 + 2
    scenes = make_scenes(1000 + rank, n_images)
    if depth:
        eng = depth_engine.DepthOrderEngine(D, max_pairs=PAIRS_PER_STEP, max_images=8, device=dev)
        eng.load_state_dict(synth.instadepth_state_dict(0))
        heads = engine.heads_for("InstaOrderNet_od", NUM_CLASSES)
    else:
        eng = engine.OrderEngine(NUM_CLASSES, D, max_pairs=PAIRS_PER_STEP, device=dev)
        eng.load_state_dict(synth.random_state_dict(0, 5, NUM_CLASSES))
        heads = engine.heads_for(ALGO, NUM_CLASSES)
    batches, mat_elems = eng.make_batches(scenes, PAIRS_PER_STEP, gmode)
    batches = batches[:n_batches]
    resident = [eng.upload_resident(b, mat_elems, gmode) for b in batches]
    input_bytes = sum(r.input_bytes for r in resident)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------
    for i in range(args.warmup):
        eng.run_resident(resident[i % len(resident)], heads, gmode)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.run_resident(resident[i % len(resident)], heads, gmode)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.gpu_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * PAIRS_PER_STEP / (ms / 1000.0)

    # ---- per-kernel events (separate pass, so the events do not perturb the number above) -------------------
    conv_ms = conv_flops = tot_ms = 0.0
    n_conv = 0
    prof_steps = 0 if depth else min(args.steps, 4)     # three handles in the InstaDepthNet engine: roofline from the step
    if not depth:
        _lib.check(eng.lib.io_net_profile(eng.net, 1))
    for i in range(prof_steps):
        eng.run_resident(resident[i % len(resident)], heads, gmode)
        torch.cuda.synchronize()
        mx = 4096
        pms = np.zeros(mx, np.float32); kind = np.zeros(mx, np.int32); fl = np.zeros(mx, np.float64)
        n = _lib.check(eng.lib.io_net_profile_read(eng.net, _lib.ptr(pms), _lib.ptr(kind), _lib.ptr(fl), None, None, mx))
        sel = (kind[:n] == 0) | (kind[:n] == 2) | (kind[:n] >= 4)
        conv_ms += float(pms[:n][sel].sum()); conv_flops += float(fl[:n][sel].sum()); n_conv += int(sel.sum())
        tot_ms += float(pms[:n].sum())
    if not depth:
        _lib.check(eng.lib.io_net_profile(eng.net, 0))
    peaks = measured_peaks()
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    if depth:
        achieved = value / world * FLOP_PER_PAIR / 1e12
    # DRAM traffic of the conv kernel per launch, from the committed ncu capture of this command (profiles/)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.isfile(tp) and not depth:
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    # ---- end to end through the public API (host numpy -> pinned -> H2D -> ... -> D2H matrices) ---------------
    per_step_scenes = []
    o = 0
    for i in range(args.warmup + args.steps):
        # the public API takes whole images: an e2e step = 17 images = 765 pairs = 3 engine batches (256+256+253)
        n_e2e = 6 if depth else 17
        per_step_scenes.append([scenes[(o + k) % len(scenes)] for k in range(n_e2e)])
        o += n_e2e
    for i in range(args.warmup):
        eng.infer_scenes(per_step_scenes[i], ALGO, "all", gmode)
    sync_all()
    h0, d0 = eng.h2d_bytes, eng.d2h_bytes
    t0 = time.perf_counter()
    pairs_e2e = 0
    for i in range(args.warmup, args.warmup + args.steps):
        r = eng.infer_scenes(per_step_scenes[i], ALGO, "all", gmode)
        pairs_e2e += sum(s.n * (s.n - 1) // 2 for s in per_step_scenes[i])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * pairs_e2e / dt
    h2d = (eng.h2d_bytes - h0) / args.steps
    d2h = (eng.d2h_bytes - d0) / args.steps

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and depth:
            v, n, cdt, threads = cpu_port_instadepth()
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind="port",
                       sample="%d pairs of one image (%.1f s): oracle port of the order branch (fp32 torch-CPU, encoder once "
                              "per image, no MiDaS decoder -- an upper bound on the reference, which recomputes both per "
                              "pair direction)" % (n, cdt))
        elif world == 1 and not args.no_cpu_baseline:
            v, n, cdt, threads = cpu_port_pairs_per_s(args.cpu_sample_pairs)
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind="port",
                       sample="%d pairs of C2 images (%.1f s): oracle port of inference.py patch path + fp32 "
                              "torch-CPU ResNet-50, batched 16 forwards" % (n, cdt))
        line = dict(
            metric=("instance pairs/s (InstaDepthNet^od order inference, resize 384^2, bf16)" if depth else
                    "instance pairs/s (InstaOrderNet^od, %s, bf16)" % ("256^2" if gmode == "patch" else "resize 384^2")),
            value=value, unit="pairs/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload="C2: synthetic COCO-val-shaped images, 10 instances -> 45 pairs/image, %d pairs per "
                                 "step, %s, %s, random-init weights" % (PAIRS_PER_STEP, "patch 256^2" if gmode == "patch" else "resize 384^2 (shipped config)", "InstaDepthNet^od order branch (ResNeXt-101 encoder layer1-3 once per image + do_net / oo_net trunks)" if depth else "InstaOrderNet^od heads [2,3]"),
                        pairs_per_step=PAIRS_PER_STEP, parallelism="images sharded over %d GPU(s), no collective" % world,
                        l2="inputs rotate over %d resident batches (%.0f MB) and each step streams >10 GB of "
                           "activations, i.e. >> 126 MB L2" % (len(resident), input_bytes / 1e6)),
            e2e=dict(value=e2e_value, unit="pairs/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                     pairs_per_step=pairs_e2e // args.steps),
            gpu_launches=launches,
            clocks=clocks,
            roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                          frac=achieved / peaks["bf16"], traffic=traffic,
                          kernel=("whole step: encoder (once per image) + two trunks, algorithmic FLOPs of this "
                                  "formulation" if depth else "conv_tc_kernel (all 53 conv layers)"), launches_timed=n_conv, avg_launch_ms=conv_ms / max(n_conv, 1),
                          conv_share_of_step=conv_ms / tot_ms if tot_ms else None, peak_source=peaks["source"],
                          step_frac=value / world * FLOP_PER_PAIR / 1e12 / peaks["bf16"]),
            cpu_baseline=cpu,
        )
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
