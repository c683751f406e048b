"""Model wrappers with the reference's names and constructor (``models.__dict__[algo](params, load_pretrain,
dist_model)``, reference models/supervised_order.py:18-95, 370-548 and models/single_stage_model.py:11-78).

Round 1 covers the inference / validation surface: ``load_state`` / ``load_pretrain`` / ``save_state`` (reference
``.pth.tar`` layout), ``switch_to``, ``model(x)``, ``set_input`` + ``forward_only`` (validation losses, reference
trainer.py:218-266) and the engine handle used by ``instaorder_b200.inference``.  ``step`` (training) raises
NotImplementedError until the backward kernels land (scope table).
"""
import os

import numpy as np
import torch

from . import _lib
from .engine import OrderEngine

__all__ = ["InstaOrderNet_o", "InstaOrderNet_d", "InstaOrderNet_od", "OrderNet"]


class _OrderModel(object):
    algo = None

    def __init__(self, params, load_pretrain=None, dist_model=False):
        self.params = params
        self.use_rgb = params.get("use_rgb", False)
        bp = params.get("backbone_param", {})
        if params.get("backbone_arch", "resnet50_cls") != "resnet50_cls":
            raise Exception("backbone_arch %r is not supported (resnet50_cls only)" % params.get("backbone_arch"))
        if bp.get("in_channels", 5) != 5:
            raise NotImplementedError("in_channels=%r: only the 5-channel (2 masks + rgb) nets are built" %
                                      bp.get("in_channels"))
        if params.get("optim", "SGD") not in ("SGD", "Adam"):
            raise Exception("No such optimizer: {}".format(params["optim"]))   # single_stage_model.py:44
        self.num_classes = bp.get("num_classes", 2)
        self.world_size = 1
        self.max_pairs = int(params.get("max_pairs", 256))
        self.device = params.get("device", "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")))
        self._engines = {}
        self._state = None
        self.phase = "eval"
        if load_pretrain is not None:
            self.load_pretrain(load_pretrain)

    # one engine (= one io_net_t workspace) per input size actually used
    def engine_for(self, input_size):
        e = self._engines.get(input_size)
        if e is None:
            if self._state is None:
                raise RuntimeError("no weights loaded: call load_state()/load_state_dict() first "
                                   "(the reference's random init gives all-tie logits)")
            e = OrderEngine(self.num_classes, input_size, self.max_pairs, self.device)
            e.load_state_dict(self._state)
            self._engines[input_size] = e
        return e

    def load_state_dict(self, sd):
        self._state = sd
        for e in self._engines.values():
            e.load_state_dict(sd)

    def load_state(self, path, Iter=None, resume=False):
        """reference models/single_stage_model.py:54-61 + utils/common_utils.py:128-149."""
        if Iter is not None:
            path = os.path.join(path, "ckpt_iter_{}.pth.tar".format(Iter))
        if not os.path.isfile(path):
            raise Exception("=> no checkpoint found at '{}'".format(path))
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        self.load_state_dict(ckpt["state_dict"])
        return ckpt["step"]

    def load_pretrain(self, load_path):
        self.load_state(load_path)

    def save_state(self, path, Iter):
        """reference models/single_stage_model.py:66-72 -- same file name and dict layout.  (No optimiser state
        exists before the training step is built; an empty dict is stored in its place.)"""
        if self._state is None:
            raise RuntimeError("no weights to save")
        path = os.path.join(path, "ckpt_iter_{}.pth.tar".format(Iter))
        sd = {(k if k.startswith("module.") else "module." + k): torch.as_tensor(np.asarray(v))
              for k, v in self._state.items()}
        torch.save({"step": Iter, "state_dict": sd, "optimizer": {}}, path)

    def switch_to(self, phase):
        if phase == "train":
            raise NotImplementedError("training mode is not built yet (round 1 = inference / validation path)")
        self.phase = phase

    # ---- model(x): eval-mode forward of an arbitrary [B,5,D,D] batch (reference resnet_cls.py:203-222) ------
    @property
    def model(self):
        return _ModelCallable(self)

    def _forward_batch(self, rgb, modal1, modal2):
        """logits [B, 2, K] fp32 (both directions) for collated fp32 NCHW tensors on the device."""
        D = int(rgb.shape[-1])
        eng = self.engine_for(D)
        dev = eng.device
        rgb = rgb.to(dev, torch.float32).contiguous()
        m1 = modal1.to(dev, torch.float32).contiguous()
        m2 = modal2.to(dev, torch.float32).contiguous()
        B = rgb.shape[0]
        out = torch.empty((B, 2, eng.k_total), dtype=torch.float32, device=dev)
        for b0 in range(0, B, eng.max_pairs):
            n = min(eng.max_pairs, B - b0)
            _lib.check(eng.lib.io_pair_pack_nchw(rgb[b0:b0 + n].data_ptr(), m1[b0:b0 + n].data_ptr(),
                                                 m2[b0:b0 + n].data_ptr(), n, D, eng.pair_tensor.data_ptr(),
                                                 _lib.stream_ptr()))
            eng.gpu_launches += 1
            eng.forward(n)
            out[b0:b0 + n].copy_(eng.logits[:n])
        return out

    # ---- validation: set_input + forward_only (reference models/supervised_order.py, per class) ----------------
    def _set_common(self, rgb, modal1, modal2):
        dev = torch.device(self.device)
        self.rgb, self.modal1, self.modal2 = rgb.to(dev), modal1.to(dev), modal2.to(dev)

    def _loss(self, logits, occ_off, class_off, class_k, occ_target, class_target, is_overlap):
        out = torch.empty(3, dtype=torch.float32, device=logits.device)
        _lib.check(_lib.lib().io_loss_forward(
            logits.data_ptr(), logits.shape[0], logits.shape[2], occ_off, class_off, class_k,
            _lib.ptr(occ_target), _lib.ptr(class_target), _lib.ptr(is_overlap),
            float(self.params.get("overlap_weight", 1.0)), float(self.params.get("distinct_weight", 1.0)),
            int(self.world_size), out.data_ptr(), _lib.stream_ptr()))
        return out

    def step(self):
        raise NotImplementedError("training step is not built yet (round 1 = inference / validation path)")


class _ModelCallable(object):
    """Stands in for ``FixModule(resnet50_cls(...))`` in eval mode: ``model.model(x)`` with x = [B,5,D,D]."""

    def __init__(self, owner):
        self.owner = owner

    def __call__(self, x):
        lg = self.owner._forward_batch(x[:, 2:5], x[:, 0:1], x[:, 1:2])[:, 0, :]
        nc = self.owner.num_classes
        if isinstance(nc, (list, tuple)):
            return lg[:, :nc[0]].contiguous(), lg[:, nc[0]:nc[0] + nc[1]].contiguous()
        return lg.contiguous()

    def eval(self):
        return self

    def state_dict(self):
        return self.owner._state


def _swap01(t):
    """order2 of set_input: 0 -> 1, 1 -> 0, everything else unchanged."""
    o = t.clone()
    o[t == 0] = 1
    o[t == 1] = 0
    return o


class InstaOrderNet_o(_OrderModel):
    algo = "InstaOrderNet_o"

    def set_input(self, rgb=None, modal1=None, modal2=None, occ_order=None):      # supervised_order.py:509-516
        self._set_common(rgb, modal1, modal2)
        self.occ_order1 = occ_order.to(self.rgb.device, torch.float32).contiguous()
        self.occ_order2 = self.occ_order1[:, [1, 0]].contiguous()

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:518-533
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        if not ret_loss:
            return {}
        return {}, {"loss": self._loss(lg, 0, -1, 0, self.occ_order1, None, None)[0]}


class InstaOrderNet_d(_OrderModel):
    algo = "InstaOrderNet_d"

    def set_input(self, rgb=None, modal1=None, modal2=None, depth_order=None, count=None, is_overlap=None):  # :383-395
        self._set_common(rgb, modal1, modal2)
        self.depth_order1 = depth_order.to(self.rgb.device, torch.int64).contiguous()
        self.depth_order2 = _swap01(self.depth_order1)
        self.count = count.to(self.rgb.device)
        self.is_overlap = is_overlap.to(self.rgb.device, torch.int64).contiguous()

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:397-411
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        if not ret_loss:
            return {}
        return {}, {"loss": self._loss(lg, -1, 0, 3, None, self.depth_order1, None)[0]}


class InstaOrderNet_od(_OrderModel):
    algo = "InstaOrderNet_od"

    def set_input(self, rgb=None, modal1=None, modal2=None, depth_order=None, count=None, is_overlap=None,
                  occ_order=None):                                                # supervised_order.py:32-48
        self._set_common(rgb, modal1, modal2)
        self.depth_order1 = depth_order.to(self.rgb.device, torch.int64).contiguous()
        self.depth_order2 = _swap01(self.depth_order1)
        self.count = count.to(self.rgb.device)
        self.is_overlap = is_overlap.to(self.rgb.device, torch.int64).contiguous()
        self.occ_order1 = occ_order.to(self.rgb.device, torch.float32).contiguous()
        self.occ_order2 = self.occ_order1[:, [1, 0]].contiguous()

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:50-81
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        out = self._loss(lg, 0, 2, 3, self.occ_order1, self.depth_order1, self.is_overlap)
        return {"loss_occ": out[1], "loss_depth": out[2]}, {"loss": out[0]}


class OrderNet(_OrderModel):
    algo = "OrderNet"

    def set_input(self, rgb=None, modal1=None, modal2=None, occ_order=None):      # supervised_order.py:451-463
        self._set_common(rgb, modal1, modal2)
        self.occ_order1 = occ_order.to(self.rgb.device, torch.int64).contiguous()
        self.occ_order2 = _swap01(self.occ_order1)

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:465-479
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        if not ret_loss:
            return {}
        k = self.num_classes
        return {}, {"loss": self._loss(lg, -1, 0, int(k), None, self.occ_order1, None)[0]}
