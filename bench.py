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
    """Roofline denominators.  ``bf16`` is the BURST cuBLAS figure: every fraction this file reports is against it
    (kernels are event-timed, the timed region lasts well under a second -- the sustained figure, taken at a 1.3 GHz
    median clock over a seconds-long loop, is reported beside it as ``*_sustained`` for context only)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        burst = j.get("bf16_tflops", 1590.0)
        return dict(bf16=burst, bf16_sustained=j.get("bf16_tflops_sustained", burst), hbm=j.get("hbm_gbs", 6650.0),
                    source="measured (MEASURED_PEAKS.json: burst bf16 %.1f TFLOP/s, HBM copy %.0f GB/s)" %
                           (burst, j.get("hbm_gbs", 6650.0)))
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a thread every 10 ms (the timed region is a
    few hundred ms; an ``nvidia-smi -lms`` child often starts too late for it), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.nvml = None
        self.samples = []      # (sm_mhz, reasons bitmask, power_w)
        self.sm_max = None
        self._stop = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((sm, rs, pw))
            except Exception:
                pass
            self._stop.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            n = self.nvml
            names = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
            sm = sorted(x[0] for x in self.samples)
            reasons = sorted({nm for _, rs, _ in self.samples for nm, bit in names if rs & bit})
            pw = [x[2] for x in self.samples if x[2] is not None]
            # every sample lies inside the timed region (start / stop bracket it): plain median
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.sm_max, reasons=reasons,
                        samples=len(sm), sm_mhz_min=min(sm) if sm else None, power_w_max=max(pw) if pw else None,
                        source="nvml")
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
                    reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")


WORKLOADS = {
    # mode -> (label, scene generator args): SURVEY.md section 8d
    "c2": ("C2: synthetic COCO-val-shaped images, 10 instances -> 45 pairs/image", None),
    "c1": ("C1: synthetic 640x480 images, 8 instances -> 28 pairs/image", dict(H=480, W=640, N=8, float_boxes=False)),
    "c3": ("C3: synthetic KINS-shaped 1242x375 images, 15 instances -> 105 pairs/image",
           dict(H=375, W=1242, N=15, wh_range=((20, 200), (20, 150)), float_boxes=False)),
}


def make_scenes(seed, n_images, workload="c2"):
    from instaorder_b200 import engine, synth
    out = []
    kw = WORKLOADS[workload][1]
    if kw is None:
        stream = synth.coco_scene_stream(seed, n_images, N=10)
    else:
        rng = np.random.RandomState(seed)
        stream = (synth.make_scene(rng, **kw) for _ in range(n_images))
    for image, masks, boxes in stream:
        out.append(engine.Scene(image, masks, engine.expand_bbox(boxes, 3.0)))
    return out


def crop_read_bytes(batch):
    """Algorithmic source bytes one gather launch reads: 5 B (rgb + two masks) per source pixel
    in crop \u2229 image, summed over the pairs of the batch."""
    tot = 0
    for (sc, pairs, crops, _, _) in batch:
        x0 = np.maximum(crops[:, 0], 0)
        y0 = np.maximum(crops[:, 1], 0)
        x1 = np.minimum(crops[:, 0] + crops[:, 2], sc.w)
        y1 = np.minimum(crops[:, 1] + crops[:, 2], sc.h)
        tot += int((np.maximum(x1 - x0, 0).astype(np.int64) * np.maximum(y1 - y0, 0)).sum()) * 5
    return tot


def time_gather(eng, resident, batches, gmode, peaks, reps=3):
    """CUDA-event time of the fused gather kernel alone (G5-G9: crop + cubic / nearest resize + normalise + bf16
    pair tensor) over the resident batches, against the measured HBM copy bandwidth."""
    import torch
    if gmode != "patch":
        return None
    for r in resident[:2]:
        eng.gather(r, r.P, gmode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for r in resident:
            eng.gather(r, r.P, gmode)
    e1.record()
    torch.cuda.synchronize()
    n = reps * len(resident)
    ms = e0.elapsed_time(e1) / n
    P = resident[0].P
    read = float(np.mean([crop_read_bytes(b) for b in batches]))
    algo = P * D * D * 5 * 2 + read                          # one 5-channel bf16 tensor per pair + the unique source bytes
    written = int(eng.lib.io_pair_tensor_bytes(P, D))      # what the kernel actually stores (8-channel pixels + border)
    gbs = algo / (ms / 1000.0) / 1e9
    return dict(bound="hbm", kernel="gather_patch_kernel", achieved=gbs, peak=peaks["hbm"], unit="GB/s",
                frac=gbs / peaks["hbm"], avg_launch_ms=ms, launches_timed=n, algorithmic_bytes_per_launch=int(algo),
                stored_bytes_per_launch=written, stored_GBps=(written + read) / (ms / 1000.0) / 1e9)


def time_metrics(dev, peaks, images=65536, n=16, reps=5):
    """the metric kernels batched over 65,536 synthetic images with N = 16 instances each
    (int64 matrices, the reference's ``np.int``): algorithmic bytes = the matrices each kernel must read once."""
    import torch
    from instaorder_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(5)
    nn_ = n * n
    order = torch.randint(0, 2, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    gt = torch.randint(-1, 2, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    dpred = torch.randint(0, 3, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    dgt = torch.randint(0, 3, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    ovl = torch.randint(0, 2, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    cnt = torch.randint(1, 4, (images * nn_,), generator=g, device=dev, dtype=torch.int64)
    off = torch.arange(images, device=dev, dtype=torch.int64) * nn_
    ns = torch.full((images,), n, device=dev, dtype=torch.int32)
    out3 = torch.empty((images, 3), dtype=torch.float64, device=dev)
    out9 = torch.empty((images, 9), dtype=torch.float64, device=dev)
    st = _lib.stream_ptr()
    res = {}

    def run(name, fn, algo_bytes):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = algo_bytes / (ms / 1000.0) / 1e9
        res[name] = dict(bound="hbm", kernel=name + "_kernel", achieved=gbs, peak=peaks["hbm"], unit="GB/s",
                         frac=gbs / peaks["hbm"], avg_launch_ms=ms, us_per_image=1000.0 * ms / images,
                         algorithmic_bytes_per_launch=int(algo_bytes), images=images, n=n)

    run("prf", lambda: _lib.check(L.io_metrics_prf(order.data_ptr(), gt.data_ptr(), off.data_ptr(), ns.data_ptr(), images,
                                                   1, out3.data_ptr(), st)), images * (2 * nn_ * 8 + 24))
    tri = n * (n - 1) // 2
    run("whdr", lambda: _lib.check(L.io_metrics_whdr(dpred.data_ptr(), dgt.data_ptr(), ovl.data_ptr(), cnt.data_ptr(),
                                                     off.data_ptr(), ns.data_ptr(), images, out9.data_ptr(), st)),
        images * (4 * tri * 8 + 72))
    return res


def host_threads():
    """Every host core this process may use (torchrun pins OMP_NUM_THREADS=1 by default)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_REF_MODEL = {}


def cpu_reference_pairs_per_s(n_pairs, threads=None, on_gpu=False):
    """The UNMODIFIED reference (``inference.infer_order_sup_occ_depth``: per-pair cv2 crops, two batch-1 fp32
    forwards, five host syncs per pair -- reference inference.py:349-436) on the host cores, imported from
    ``/root/reference`` or, on the GPU box, from the archive oracle/build_ref.py packed into ``oracle/_ref/``.
    Returns None when neither exists (the caller falls back to the oracle port run the same batch-1 way).
    ``on_gpu``: leave the reference's own ``.cuda()`` calls alone -- its stock PyTorch-eager GPU path (cuDNN / cuBLAS,
    fp32, batch 1) on this same B200, reported beside the CPU number for context."""
    from oracle import ref_shim
    if not ref_shim.available():
        return None
    import torch
    if on_gpu and not torch.cuda.is_available():
        return None
    if on_gpu:
        return _reference_run(n_pairs, threads, "gpu")
    with ref_shim.cpu_only():
        return _reference_run(n_pairs, threads, "cpu")


def _reference_run(n_pairs, threads, where):
    import torch
    from instaorder_b200 import synth
    from oracle import gen_golden, oracle as O, ref_shim
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    ns = ref_shim.load()
    if where not in _REF_MODEL:
        _REF_MODEL[where] = gen_golden.make_reference_model(ns, ALGO, NUM_CLASSES,
                                                            synth.random_state_dict(0, 5, NUM_CLASSES))
    model = _REF_MODEL[where]
    n_img = max(1, (n_pairs + 44) // 45)
    scenes = list(synth.coco_scene_stream(99, n_img, N=10))
    image, masks, boxes = scenes[0]
    bexp = O.expand_bbox(boxes, 3.0)
    ns.inference.infer_order_sup_occ_depth(model, image, masks[:2], bexp[:2], "all", ALGO, "patch", D, "")   # warm-up
    done, dt = 0, 0.0
    for (image, masks, boxes) in scenes:
        left = n_pairs - done
        if left <= 0:
            break
        k = 2
        while k * (k - 1) // 2 < left and k < 10:
            k += 1
        bexp = O.expand_bbox(boxes, 3.0)
        if where == "gpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ns.inference.infer_order_sup_occ_depth(model, image, masks[:k], bexp[:k], "all", ALGO, "patch", D, "")
        if where == "gpu":
            torch.cuda.synchronize()
        dt += time.perf_counter() - t0
        done += k * (k - 1) // 2
    return done / dt, done, dt, torch.get_num_threads()


def cpu_port_pairs_per_s(n_pairs, threads=None, batch1=False):
    """The oracle port (= the reference algorithm: per-pair cv2-equivalent crops, fp32 torch-CPU ResNet-50 on both
    directions, decisions) timed on the host cores for a bounded sample of the same workload.  ``batch1``: one forward
    per network input as the reference does; otherwise 16 forwards per batch (the best case for a CPU)."""
    import torch
    from instaorder_b200 import synth
    from oracle import oracle as O
    threads = threads or host_threads()
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
        r = O.infer_order(sd, image, masks[:k], bexp[:k], "all", ALGO, "patch", D, chunk=1 if batch1 else 8,
                          batch1=batch1)
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
    steps, warmup = args.steps, args.warmup
    sample_pairs = 45                                           # one whole C2 image per step: 10 instances -> 45 pairs (~1.6 s)
    vals = []
    kind = "reference"
    for s in range(warmup + steps):
        r = cpu_reference_pairs_per_s(sample_pairs)
        if r is None:                                           # no reference tree / archive: the port, run batch-1
            kind = "port"
            r = cpu_port_pairs_per_s(sample_pairs, batch1=True)
        v, n, dt, threads = r
        if s >= warmup:
            vals.append((n, dt))
    pairs = sum(n for n, _ in vals)
    secs = sum(dt for _, dt in vals)
    value = pairs / secs
    how = ("the UNMODIFIED reference's inference.infer_order_sup_occ_depth (per-pair cv2 crops + two batch-1 fp32 "
           "torch-CPU forwards, oracle/_ref archive)" if kind == "reference" else
           "fp32 torch-CPU oracle port of the reference's inference.py patch path, two batch-1 forwards per pair")
    # same metric / config.workload strings as the b200 arm (the driver pairs the two lines); the reference arm computes in
    # fp32 (dtype) and its step is a bounded sample of the workload (config.sample)
    line = dict(impl="reference", metric="instance pairs/s (InstaOrderNet^od, 256^2, bf16)", value=value, unit="pairs/s",
                n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=1000.0 * secs / max(len(vals), 1),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="C2: synthetic COCO-val-shaped images, 10 instances -> 45 pairs/image, %d pairs per "
                                     "step, patch 256^2, InstaOrderNet^od heads [2,3], random-init weights" % PAIRS_PER_STEP,
                            sample="bounded sample of %d pairs per step on the host cores: %s" % (sample_pairs, how)),
                cpu_baseline=dict(value=value, unit="pairs/s", cores=threads, kind=kind,
                                  sample="%d pairs/step x %d steps: %s" % (sample_pairs, steps, how)),
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
    line = measure_train(args, args.steps, args.warmup)
    if rank == 0:
        emit(line)
    import torch.distributed as dist
    if world > 1:
        dist.destroy_process_group()


def init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    return world, rank, local, dev


def measure_train(args, steps, warmup, with_e2e=True, with_cpu=True):
    """One training measurement (all ranks call it; the returned line is complete on rank 0)."""
    import torch
    import torch.distributed as dist
    from instaorder_b200 import _lib, models, synth
    world, rank, local, dev = init_dist()
    B = args.train_batch
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
        host.append(dict(rgb=torch.randn((B, 3, 256, 256), generator=g).pin_memory(),
                         modal1=(torch.rand((B, 1, 256, 256), generator=g) > 0.7).float().pin_memory(),
                         modal2=(torch.rand((B, 1, 256, 256), generator=g) > 0.7).float().pin_memory(),
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

    for i in range(warmup):
        step(resident[i % n_batches])
    sync_all()
    eng = model._trainer
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
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
    value = world * steps * B / (ms / 1000.0)
    e2e = None
    if with_e2e:
        # end to end: pinned host batch -> H2D -> step -> loss read back, every step
        for i in range(2):
            float(step(host[i % n_batches])[1]["loss"])
        sync_all()
        t0 = time.perf_counter()
        for i in range(steps):
            float(step(host[i % n_batches])[1]["loss"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = dict(value=world * steps * B / dt, unit="pairs/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4)
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
    cpu = None
    if rank == 0 and world == 1 and with_cpu and not args.no_cpu_baseline:
        v, npairs, cdt, threads = cpu_train_pairs_per_s(2, 256)
        cpu = dict(value=v, unit="pairs/s", cores=threads, kind="port",
                   sample="one step() on 2 pairs (%.1f s): training oracle, fp32 torch-CPU autograd" % cdt)
    achieved = tc_fl / (tc_ms / 1000.0) / 1e12 if tc_ms > 0 else 0.0
    step_tf = value / world * 3 * 21.764e9 / 1e12
    return dict(
        metric="training pairs/s (InstaOrderNet^od step: fwd + bwd + all-reduce + SGD, 256^2, bf16)",
        value=value, unit="pairs/s", n_gpus=world, steps=steps, warmup=warmup,
        ms_per_step=ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
        data="synthetic",
        config=dict(workload="C4: InstaOrderNet^od training step, %d synthetic pairs per GPU per step at 256^2, SGD "
                             "lr 1e-4 momentum 0.9 wd 1e-4, random-init weights" % B,
                    pairs_per_step=B * world, parallelism="data parallel over %d GPU(s): the flat fp32 gradient buffer "
                    "(94 MB) is all-reduced over NCCL every step" % world,
                    l2="each step streams > 10 GB of saved activations and gradients, i.e. >> 126 MB L2"),
        e2e=e2e, gpu_launches=launches, clocks=clocks,
        roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                      frac=achieved / peaks["bf16"], frac_sustained=achieved / peaks["bf16_sustained"], traffic=None,
                      kernel="conv_tc / conv_tn (forward + data gradient) + wgrad_kernel",
                      tensor_share_of_step=tc_ms / float(pms[:n].sum()), peak_source=peaks["source"],
                      elementwise=dict(bound="hbm", achieved=ew_by / (ew_ms / 1000.0) / 1e9 if ew_ms else 0.0,
                                       peak=peaks["hbm"], unit="GB/s",
                                       frac=ew_by / (ew_ms / 1000.0) / 1e9 / peaks["hbm"] if ew_ms else 0.0,
                                       share_of_step=ew_ms / float(pms[:n].sum())),
                      step_frac=step_tf / peaks["bf16"], step_frac_sustained=step_tf / peaks["bf16_sustained"]),
        cpu_baseline=cpu)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-pairs", type=int, default=225)
    ap.add_argument("--mode", default="patch256", choices=["patch256", "resize384", "instadepth384", "c1_o256",
                                                           "c3_ordernet256"],
                    help="patch256 = the BASELINE.json metric (default); resize384 = the shipped InstaOrderNet^od "
                         "config (whole image -> 384^2); c1_o256 / c3_ordernet256 = BASELINE configs 1 and 3 "
                         "(InstaOrderNet^o on 640x480 / 8 instances; OrderNet on KINS-shaped 1242x375 / 15 instances)")
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

    wl, ncs = "c2", NUM_CLASSES
    if args.mode == "c1_o256":
        wl, ALGO, ncs = "c1", "InstaOrderNet_o", 2
    if args.mode == "c3_ordernet256":
        wl, ALGO, ncs = "c3", "OrderNet", 3
    world, rank, local, dev = init_dist()

    # ---- workload: every rank owns its own images (weak scaling: images are sharded, no data-path collective)
    n_batches = 8                                    # distinct resident batches, rotated so inputs never sit in L2
    ppi = {"c2": 45, "c1": 28, "c3": 105}[wl]
    n_images = (n_batches * PAIRS_PER_STEP) // ppi + 2
    scenes = make_scenes(1000 + rank, n_images, wl)
    if depth:
        eng = depth_engine.DepthOrderEngine(D, max_pairs=PAIRS_PER_STEP, max_images=8, device=dev)
        eng.load_state_dict(synth.instadepth_state_dict(0))
        heads = engine.heads_for("InstaOrderNet_od", NUM_CLASSES)
    else:
        eng = engine.OrderEngine(ncs, D, max_pairs=PAIRS_PER_STEP, device=dev)
        eng.load_state_dict(synth.random_state_dict(0, 5, ncs))
        heads = engine.heads_for(ALGO, ncs)
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
    # ---- the HBM-bound kernels on their own: fused gather, batched metrics -------------
    gather_roof = time_gather(eng, resident, batches, gmode, peaks) if not depth else None
    metrics_roof = time_metrics(dev, peaks) if (rank == 0 and args.mode == "patch256") else None
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
        n_e2e = 6 if depth else {"c2": 17, "c1": 27, "c3": 7}[wl]
        per_step_scenes.append([scenes[(o + k) % len(scenes)] for k in range(n_e2e)])
        o += n_e2e
    for i in range(args.warmup):
        eng.infer_scenes(per_step_scenes[i], ALGO, "all", gmode)
    sync_all()
    # (i) one blocking call per step (`infer_scenes`: returns when the step's matrices are on the host)
    t0 = time.perf_counter()
    pairs_e2e = 0
    for i in range(args.warmup, args.warmup + args.steps):
        r = eng.infer_scenes(per_step_scenes[i], ALGO, "all", gmode)
        pairs_e2e += sum(s.n * (s.n - 1) // 2 for s in per_step_scenes[i])
    torch.cuda.synchronize()
    dt_call = time.perf_counter() - t0
    # (ii) the streaming form of the same API (`infer_stream`: same calls, same results, two in flight -- the packing and
    # H2D of step k + 1 overlap the kernels of step k); every step's H2D and D2H is still inside the timed region
    stream_ok = hasattr(eng, "infer_stream") and not depth
    sync_all()
    h0, d0 = eng.h2d_bytes, eng.d2h_bytes
    t0 = time.perf_counter()
    if stream_ok:
        n_out = 0
        for r in eng.infer_stream(per_step_scenes[args.warmup:args.warmup + args.steps], ALGO, "all", gmode, depth=2):
            n_out += len(r)
    else:
        for i in range(args.warmup, args.warmup + args.steps):
            r = eng.infer_scenes(per_step_scenes[i], ALGO, "all", gmode)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, dt_call], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_call = float(t[0].item()), float(t[1].item())
    e2e_value = world * pairs_e2e / dt
    e2e_per_call = world * pairs_e2e / dt_call
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
        elif world == 1 and not args.no_cpu_baseline and args.mode == "patch256":
            # (i) the faithful baseline: the reference as written (two batch-1 forwards per pair); (ii) beside it the
            # best case for a CPU: the same algorithm with 16 forwards per batch (the oracle port)
            r = cpu_reference_pairs_per_s(max(args.cpu_sample_pairs, 300))     # ~ 10 s on 16 host cores
            kind = "reference"
            if r is None:
                kind, r = "port", cpu_port_pairs_per_s(args.cpu_sample_pairs // 3, batch1=True)
            v, n, cdt, threads = r
            vb, nb, cdtb, _ = cpu_port_pairs_per_s(args.cpu_sample_pairs // 2)
            rg = cpu_reference_pairs_per_s(90, on_gpu=True) if kind == "reference" else None
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind=kind,
                       sample="%d pairs of C2 images (%.1f s): %s, two batch-1 fp32 torch-CPU forwards per pair as in "
                              "inference.py:140-169" % (n, cdt, "the UNMODIFIED reference (oracle/_ref archive)"
                                                        if kind == "reference" else "oracle port of the reference"),
                       batched_port=dict(value=vb, unit="pairs/s", kind="port",
                                         sample="%d pairs (%.1f s): oracle port, 16 forwards per batch (best case for "
                                                "the host cores)" % (nb, cdtb)),
                       # context, not a CPU number: the unmodified reference's stock PyTorch-eager path (its own
                       # .cuda() calls, cuDNN fp32, batch 1, five .item() syncs per pair) on this same B200
                       reference_on_this_gpu=None if rg is None else dict(
                           value=rg[0], unit="pairs/s", kind="reference",
                           sample="%d pairs (%.1f s): inference.infer_order_sup_occ_depth as written, PyTorch eager on "
                                  "cuda:0" % (rg[1], rg[2])))
        line = dict(
            metric=("instance pairs/s (InstaDepthNet^od order inference, resize 384^2, bf16)" if depth else
                    "instance pairs/s (%s, %s, bf16)" % ({"InstaOrderNet_od": "InstaOrderNet^od", "InstaOrderNet_o":
                                                          "InstaOrderNet^o", "OrderNet": "OrderNet"}[ALGO],
                                                         "256^2" if gmode == "patch" else "resize 384^2")),
            value=value, unit="pairs/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload="%s, %d pairs per step, %s, %s, random-init weights" % (
                WORKLOADS[wl][0], PAIRS_PER_STEP, "patch 256^2" if gmode == "patch" else "resize 384^2 (shipped config)",
                "InstaDepthNet^od order branch (ResNeXt-101 encoder layer1-3 once per image + do_net / oo_net trunks)"
                if depth else {"InstaOrderNet_od": "InstaOrderNet^od heads [2,3]", "InstaOrderNet_o": "InstaOrderNet^o head [2]",
                               "OrderNet": "OrderNet head [3]"}[ALGO]),
                        pairs_per_step=PAIRS_PER_STEP, parallelism="images sharded over %d GPU(s), no collective" % world,
                        l2="inputs rotate over %d resident batches (%.0f MB) and each step streams >10 GB of "
                           "activations, i.e. >> 126 MB L2" % (len(resident), input_bytes / 1e6)),
            e2e=dict(value=e2e_value, unit="pairs/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                     pairs_per_step=pairs_e2e // args.steps,
                     api="OrderEngine.infer_stream (two calls in flight)" if stream_ok else "infer_scenes",
                     per_call=dict(value=e2e_per_call, unit="pairs/s", api="OrderEngine.infer_scenes (one blocking call "
                                   "per step: nothing overlaps the first batch's packing or the final D2H)")),
            gpu_launches=launches,
            clocks=clocks,
            # frac: the dominant kernel family (all convolution launches, CUDA events inside the library) against the
            # BURST measured bf16 peak; step_frac: the whole step, pairs/s x 21.764 GFLOP / burst peak (BASELINE.md
            # section 2) -- the headline fraction; *_sustained: the same against the seconds-long cuBLAS figure
            roofline=dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                          frac=achieved / peaks["bf16"], frac_sustained=achieved / peaks["bf16_sustained"],
                          traffic=traffic,
                          kernel=("whole step: encoder (once per image) + two trunks, algorithmic FLOPs of this "
                                  "formulation" if depth else
                                  "conv_tc (single CTA / CTA pair) / conv_tn / conv_row3 / conv_fused / stem_pool kernels (all 53 conv "
                                  "layers)"),
                          launches_timed=n_conv, avg_launch_ms=conv_ms / max(n_conv, 1),
                          conv_share_of_step=conv_ms / tot_ms if tot_ms else None, peak_source=peaks["source"],
                          step_tflops=value / world * FLOP_PER_PAIR / 1e12,
                          step_frac=value / world * FLOP_PER_PAIR / 1e12 / peaks["bf16"],
                          step_frac_sustained=value / world * FLOP_PER_PAIR / 1e12 / peaks["bf16_sustained"],
                          gather=gather_roof, metrics=metrics_roof),
            cpu_baseline=cpu,
        )
    # ---- BASELINE config C4 beside the headline: the training step (the path with a real collective), so that the
    # driver's 1 -> 8 GPU runs of this file also carry a training scaling curve.  INSTAORDER_BENCH_TRAIN=0 skips it.
    if args.mode == "patch256" and os.environ.get("INSTAORDER_BENCH_TRAIN", "1") != "0":
        del resident
        eng = None
        torch.cuda.empty_cache()
        tr = measure_train(args, steps=min(args.steps, 20), warmup=max(3, min(args.warmup, 5)), with_e2e=False,
                           with_cpu=False)
        if rank == 0:
            line["training"] = {k: tr[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                                                   "scaling", "gpu_launches", "clocks", "roofline", "config")}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
