"""Model wrappers with the reference's names and constructor (``models.__dict__[algo](params, load_pretrain,
dist_model)``, reference models/supervised_order.py:18-95, 370-548 and models/single_stage_model.py:11-78).

Round 1 covers the inference surface: ``load_state`` / ``load_pretrain`` (reference ``.pth.tar`` layout),
``switch_to`` and the engine handle used by ``instaorder_b200.inference``.  ``set_input`` / ``step`` /
``forward_only`` (training) raise NotImplementedError until the backward kernels land (scope table).
"""
import os

import torch

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

    def switch_to(self, phase):
        if phase == "train":
            raise NotImplementedError("training mode is not built yet (round 1 = inference path)")
        self.phase = phase

    def set_input(self, *a, **k):
        raise NotImplementedError("training step is not built yet (round 1 = inference path)")

    step = forward_only = set_input


class InstaOrderNet_o(_OrderModel):
    algo = "InstaOrderNet_o"


class InstaOrderNet_d(_OrderModel):
    algo = "InstaOrderNet_d"


class InstaOrderNet_od(_OrderModel):
    algo = "InstaOrderNet_od"


class OrderNet(_OrderModel):
    algo = "OrderNet"
