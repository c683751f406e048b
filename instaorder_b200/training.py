"""Host side of the training step: owns the flat fp32 parameter / gradient / optimiser buffers (torch tensors on the
device), maps them to the reference's ``state_dict`` names, and drives ``io_train_*`` of the C ABI.

Replaces (reference): ``SingleStageModel.__init__`` optimiser construction (models/single_stage_model.py:34-42),
the body of ``step()`` (models/supervised_order.py:83-95 etc.), ``utils.average_gradients``
(utils/distributed_utils.py:27-31: 163 per-tensor all-reduces -> ONE all-reduce of the flat gradient buffer) and the
optimiser part of ``save_state`` / ``load_state`` (single_stage_model.py:54-72, utils/common_utils.py:128-149).
PyTorch provides device memory, streams and ``torch.distributed`` (NCCL) only; there is no CPU fallback.
"""
import ctypes as C
import collections
import os

import numpy as np
import torch

from . import _lib


def all_reduce_sum_(flat):
    """utils.average_gradients (utils/distributed_utils.py:27-31) on the flat buffer: the reference all-reduces every
    ``param.grad`` with SUM (the loss is already divided by world_size, models/supervised_order.py:78); here it is
    ONE collective over all 23.5 M gradients.  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
    return flat


def broadcast_(flat, src=0):
    """DistModule.broadcast_params (utils/distributed_utils.py:17-24) on a flat buffer."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src)
    return flat


class FlatOptim(object):
    """Stands in for ``torch.optim.SGD`` / ``Adam`` of the reference wrapper: ``param_groups[0]['lr']`` is what
    ``utils.StepLRScheduler`` writes (utils/scheduler.py:77-80), ``state_dict()`` / ``load_state_dict()`` keep the
    reference checkpoint's optimiser layout (one entry per parameter in ``model.parameters()`` order)."""

    def __init__(self, kind, lr, weight_decay=0.0, beta1=0.9):
        self.engine = None
        self.kind = kind
        if kind == "SGD":
            self.param_groups = [dict(lr=lr, initial_lr=lr, momentum=0.9, dampening=0, weight_decay=weight_decay,
                                      nesterov=False)]
        else:
            self.param_groups = [dict(lr=lr, initial_lr=lr, betas=(beta1, 0.999), eps=1e-8, weight_decay=0,
                                      amsgrad=False)]
        self.steps = 0
        self.buf = self.buf2 = None
        self._pending = None      # optimiser state loaded (resume) before the engine exists

    def attach(self, engine):
        """Binds the optimiser to a TrainEngine (allocates the flat momentum / Adam buffers on its device)."""
        self.engine = engine
        n, dev = engine.n_params, engine.device
        self.buf = torch.zeros(n, dtype=torch.float32, device=dev)               # momentum / exp_avg
        self.buf2 = torch.zeros(n, dtype=torch.float32, device=dev) if self.kind == "Adam" else None   # exp_avg_sq
        if self._pending is not None:
            pend, self._pending = self._pending, None
            self.load_state_dict(pend)

    def zero_grad(self):
        pass   # the gradient buffer is zeroed inside io_train_forward_backward

    def step(self):
        e = self.engine
        g = self.param_groups[0]
        if self.kind == "SGD":
            _lib.check(e.lib.io_train_sgd_step(e.handle, self.buf.data_ptr(), float(g["lr"]), float(g["momentum"]),
                                               float(g["weight_decay"]), int(self.steps == 0), _lib.stream_ptr()))
        else:
            _lib.check(e.lib.io_train_adam_step(e.handle, self.buf.data_ptr(), self.buf2.data_ptr(), float(g["lr"]),
                                                float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                                self.steps + 1, _lib.stream_ptr()))
        self.steps += 1
        e.gpu_launches += 2

    def state_dict(self):
        e = self.engine
        state = {}
        if e is None:
            return self._pending if self._pending is not None else {"state": {}, "param_groups": self.param_groups}
        if self.steps > 0:
            bufs = e.export_flat(self.buf, params_only=True)
            bufs2 = e.export_flat(self.buf2, params_only=True) if self.buf2 is not None else None
            for i, name in enumerate(e.param_names):
                if self.kind == "SGD":
                    state[i] = {"momentum_buffer": bufs[name]}
                else:
                    state[i] = {"step": torch.tensor(float(self.steps)), "exp_avg": bufs[name],
                                "exp_avg_sq": bufs2[name]}
        pg = dict(self.param_groups[0])
        pg["params"] = list(range(len(e.param_names)))
        return {"state": state, "param_groups": [pg]}

    def load_state_dict(self, sd):
        e = self.engine
        if e is None:
            self._pending = sd
            if sd.get("param_groups"):
                for k, v in sd["param_groups"][0].items():
                    if k != "params":
                        self.param_groups[0][k] = v
            return
        st = sd.get("state", {})
        if len(st):
            key = "momentum_buffer" if self.kind == "SGD" else "exp_avg"
            e.import_flat(self.buf, {name: st[i][key] for i, name in enumerate(e.param_names) if i in st})
            if self.kind == "Adam":
                e.import_flat(self.buf2, {name: st[i]["exp_avg_sq"] for i, name in enumerate(e.param_names) if i in st})
                self.steps = int(float(st[0]["step"]))
            else:
                self.steps = 1
        if sd.get("param_groups"):
            for k, v in sd["param_groups"][0].items():
                if k != "params":
                    self.param_groups[0][k] = v


class TrainEngine(object):
    """One model replica on one GPU: ``io_train_t`` + the flat buffers."""

    def __init__(self, num_classes, input_size, batch_pairs, device="cuda:0"):
        self.lib = _lib.lib()      # raises ImportError if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("TrainEngine needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        ncs = list(num_classes) if isinstance(num_classes, (list, tuple)) else [int(num_classes)]
        self.num_classes = ncs
        self.k_total = int(sum(ncs))
        self.input_size = int(input_size)
        self.batch_pairs = int(batch_pairs)
        h = C.c_void_p()
        arr = (C.c_int32 * len(ncs))(*ncs)
        _lib.check(self.lib.io_train_create(arr, len(ncs), self.input_size, self.batch_pairs, C.byref(h)))
        self.handle = h
        self.n_params = int(self.lib.io_train_param_count(h))
        self.n_stats = int(self.lib.io_train_stat_count(h))
        self.segments = collections.OrderedDict()   # name -> (buffer, offset, dims)
        name = C.create_string_buffer(128)
        buf = C.c_int32()
        off = C.c_int64()
        dims = (C.c_int32 * 4)()
        for i in range(self.lib.io_train_num_segments(h)):
            _lib.check(self.lib.io_train_segment(h, i, name, 128, C.byref(buf), C.byref(off), dims))
            self.segments[name.value.decode()] = (int(buf.value), int(off.value), tuple(int(d) for d in dims))
        # trainable tensors in the reference's model.parameters() order (module registration order)
        from .synth import resnet50_layout
        nc = ncs if len(ncs) == 2 else ncs[0]
        self.param_names = [k for k, _ in resnet50_layout(5, nc)
                            if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        assert all(k in self.segments for k in self.param_names)
        dev = self.device
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(self.n_stats, dtype=torch.float32, device=dev)
        self.num_batches_tracked = 0
        _lib.check(self.lib.io_train_bind(h, self.params.data_ptr(), self.grads.data_ptr(), self.stats.data_ptr()))
        self.pair_tensor = torch.zeros(int(self.lib.io_pair_tensor_bytes(self.batch_pairs, self.input_size)),
                                       dtype=torch.uint8, device=dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.gpu_launches = 0
        self.loaded = False
        self._buckets = None
        self._comm_stream = None
        self.overlap_allreduce = os.environ.get("INSTAORDER_ALLREDUCE_OVERLAP", "1") != "0"

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.io_train_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- flat buffer <-> reference-layout tensors ---------------------------------------------------------------
    def _view(self, flat, name):
        _, off, dims = self.segments[name]
        if dims[1] and dims[2]:                 # conv weight, stored [cout, kh, kw, cin]
            n = dims[0] * dims[1] * dims[2] * dims[3]
            return flat[off:off + n].view(dims[0], dims[1], dims[2], dims[3]), True
        if dims[1]:                             # FC weight [rows, cols]
            return flat[off:off + dims[0] * dims[1]].view(dims[0], dims[1]), False
        return flat[off:off + dims[0]], False

    def export_flat(self, flat, params_only=False, stats=None):
        """name -> CPU tensor in the reference layout ([cout, cin, kh, kw] convolution weights)."""
        out = collections.OrderedDict()
        for name, (b, _, _) in self.segments.items():
            src = flat if b == 0 else stats
            if src is None or (params_only and b != 0):
                continue
            v, is_conv = self._view(src, name)
            out[name] = (v.permute(0, 3, 1, 2) if is_conv else v).contiguous().cpu().clone()
        return out

    def import_flat(self, flat, tensors, stats=None):
        for name, t in tensors.items():
            if name not in self.segments:
                continue
            b = self.segments[name][0]
            dst_flat = flat if b == 0 else stats
            if dst_flat is None:
                continue
            v, is_conv = self._view(dst_flat, name)
            t = torch.as_tensor(np.asarray(t) if not isinstance(t, torch.Tensor) else t).to(self.device, torch.float32)
            if is_conv:
                t = t.permute(0, 2, 3, 1)
            if tuple(t.shape) != tuple(v.shape):
                raise ValueError("%s: checkpoint tensor %r does not match %r" % (name, tuple(t.shape), tuple(v.shape)))
            v.copy_(t)

    def load_state_dict(self, sd):
        """Reference-layout state_dict (with or without the ``module.`` prefix) -> flat buffers."""
        clean = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        missing = [k for k in self.segments if k not in clean]
        if missing:
            raise KeyError("missing keys in the checkpoint: %s ..." % missing[:4])
        self.import_flat(self.params, {k: v for k, v in clean.items() if self.segments.get(k, (1,))[0] == 0})
        self.import_flat(None, {k: v for k, v in clean.items() if self.segments.get(k, (0,))[0] == 1},
                         stats=self.stats)
        nbt = clean.get("bn1.num_batches_tracked")
        self.num_batches_tracked = int(np.asarray(nbt)) if nbt is not None else 0
        _lib.check(self.lib.io_train_sync_weights(self.handle, _lib.stream_ptr()))
        self.loaded = True

    def state_dict(self, prefix="module."):
        """Reference ``state_dict`` (same keys, order and layouts; ``num_batches_tracked`` included)."""
        p = self.export_flat(self.params, stats=self.stats)
        from .synth import resnet50_layout
        nc = self.num_classes if len(self.num_classes) == 2 else self.num_classes[0]
        out = collections.OrderedDict()
        for k, _ in resnet50_layout(5, nc):
            if k.endswith("num_batches_tracked"):
                out[prefix + k] = torch.tensor(self.num_batches_tracked, dtype=torch.int64)
            else:
                out[prefix + k] = p[k]
        return out

    # ---- the step ---------------------------------------------------------------------------------------------
    def pack_inputs(self, rgb, modal1, modal2):
        """Collated fp32 NCHW tensors (supervised_order.py:34-36) -> the pair tensor (torch.cat of :84, in bf16)."""
        B, D = int(rgb.shape[0]), int(rgb.shape[-1])
        if B != self.batch_pairs or D != self.input_size:
            raise ValueError("batch [%d, %d^2] does not match the engine ([%d, %d^2])" %
                             (B, D, self.batch_pairs, self.input_size))
        dev = self.device
        rgb = rgb.to(dev, torch.float32).contiguous()
        m1 = modal1.to(dev, torch.float32).contiguous()
        m2 = modal2.to(dev, torch.float32).contiguous()
        _lib.check(self.lib.io_pair_pack_nchw(rgb.data_ptr(), m1.data_ptr(), m2.data_ptr(), B, D,
                                              self.pair_tensor.data_ptr(), _lib.stream_ptr()))
        self.gpu_launches += 1

    def forward_backward(self, occ_off, class_off, class_k, occ_target, class_target, is_overlap, overlap_w,
                         distinct_w, world_size, backward=True):
        if not self.loaded:
            raise RuntimeError("no weights loaded: call load_state_dict() first")
        _lib.check(self.lib.io_train_forward_backward(
            self.handle, self.pair_tensor.data_ptr(), occ_off, class_off, class_k, _lib.ptr(occ_target),
            _lib.ptr(class_target), _lib.ptr(is_overlap), float(overlap_w), float(distinct_w), int(world_size),
            self.losses.data_ptr(), int(backward), _lib.stream_ptr()))
        self.gpu_launches += int(self.lib.io_train_last_launches(self.handle))
        self.num_batches_tracked += 2        # two train-mode forward passes per step
        return self.losses

    def logits(self):
        """[2, B, K] fp32 logits of the last forward ([direction][pair]) as a device tensor (copy)."""
        out = torch.empty((2, self.batch_pairs, self.k_total), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.io_train_read_logits(self.handle, out.data_ptr(), _lib.stream_ptr()))
        return out

    def activation(self, conv_name, which):
        """Saved tensor of the last forward as a flat bf16 device tensor (which = 0 raw conv output, 1 activation)."""
        n = C.c_int64()
        _lib.check(self.lib.io_train_read_activation(self.handle, conv_name.encode(), which, None, C.byref(n), None))
        out = torch.empty(n.value, dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.io_train_read_activation(self.handle, conv_name.encode(), which, out.data_ptr(),
                                                     C.byref(n), _lib.stream_ptr()))
        return out

    def buckets(self):
        """[(begin, end)] element ranges of the flat gradient buffer in the order backward completes them:
        (layer4 + heads), layer3, layer2, (stem + layer1)."""
        if self._buckets is None:
            out = []
            b, e = C.c_int64(), C.c_int64()
            for k in range(self.lib.io_train_num_buckets(self.handle)):
                _lib.check(self.lib.io_train_bucket(self.handle, k, C.byref(b), C.byref(e)))
                out.append((int(b.value), int(e.value)))
            self._buckets = out
        return self._buckets

    def all_reduce_grads(self):
        """utils.average_gradients (utils/distributed_utils.py:27-31: 163 blocking per-parameter all-reduces after
        backward; SUM, the loss is pre-divided by world_size).  Here: FOUR all-reduces of contiguous ranges of the flat
        buffer, each issued on a communication stream that waits (device side) only for ITS range of the backward pass
        -- the 60 MB of layer4 + heads travel over NVLink while layer3 .. stem are still being differentiated; the
        optimiser (caller's stream) waits for all four.  INSTAORDER_ALLREDUCE_OVERLAP=0: one call after backward."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if not self.overlap_allreduce or not self.grads.is_cuda:
            all_reduce_sum_(self.grads)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        comm = self._comm_stream
        works = []
        for k, (b, e) in enumerate(self.buckets()):
            _lib.check(self.lib.io_train_wait_bucket(self.handle, k, comm.cuda_stream))
            with torch.cuda.stream(comm):
                works.append(dist.all_reduce(self.grads[b:e], async_op=True))
        for w in works:
            w.wait()          # the caller's stream waits for the NCCL stream; no host synchronisation

    def broadcast_params(self):
        """DistModule.broadcast_params (utils/distributed_utils.py:17-24): rank 0's parameters to everyone."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            broadcast_(self.params, 0)
            broadcast_(self.stats, 0)
            _lib.check(self.lib.io_train_sync_weights(self.handle, _lib.stream_ptr()))
