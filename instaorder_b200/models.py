"""Model wrappers with the reference's names and constructor (``models.__dict__[algo](params, load_pretrain,
dist_model)``, reference models/supervised_order.py:18-95, 370-548 and models/single_stage_model.py:11-78).

Covers ``load_state`` / ``load_pretrain`` / ``save_state`` (reference ``.pth.tar`` layout incl. the optimiser state),
``switch_to``, ``model(x)``, ``set_input`` + ``forward_only`` (validation losses, reference trainer.py:218-266),
``step()`` (training: train-mode BN forward of both directions, loss, backward, ONE flat gradient all-reduce,
fused SGD / Adam) and the engine handle used by ``instaorder_b200.inference``.
"""
import os

import numpy as np
import torch

from . import _lib
from .engine import OrderEngine
from .init import reference_init_state_dict
from .training import FlatOptim, TrainEngine

__all__ = ["InstaOrderNet_o", "InstaOrderNet_d", "InstaOrderNet_od", "OrderNet", "InstaDepthNet_od", "InstaDepthNet_d", "MidasNet"]


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
        self.dist_model = bool(dist_model)
        if dist_model:                                   # single_stage_model.py:28-30
            import torch.distributed as dist
            self.world_size = dist.get_world_size()
        if params.get("optim", "SGD") == "SGD":          # single_stage_model.py:34-42
            self.optim = FlatOptim("SGD", params.get("lr", 1e-4), weight_decay=params.get("weight_decay", 0.0))
        else:
            self.optim = FlatOptim("Adam", params.get("lr", 1e-4), beta1=params.get("beta1", 0.9))
        self._trainer = None        # TrainEngine, created by the first step() (batch size / input size known then)
        self._train_dirty = False   # the TrainEngine holds newer weights than self._state / the eval engines
        self.max_pairs = int(params.get("max_pairs", 256))
        self.device = params.get("device", "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")))
        self._engines = {}
        # single_stage_model.py:24-25: the constructor always draws the reference's random initialisation from
        # torch's global CPU generator (kaiming, then init_weights 'xavier' gain 0.02) -- bit-identical state_dict
        # after the same torch.manual_seed, so Trainer(args).run() starts from scratch as every train.sh does
        self._state = reference_init_state_dict(self.num_classes, bp.get("in_channels", 5))
        self.phase = "eval"
        if load_pretrain is not None:
            self.load_pretrain(load_pretrain)

    # one engine (= one io_net_t workspace) per input size actually used
    def engine_for(self, input_size):
        e = self._engines.get(input_size)
        if e is None:
            e = OrderEngine(self.num_classes, input_size, self.max_pairs, self.device)
            e.load_state_dict(self._state)
            self._engines[input_size] = e
        return e

    def engine_for_orig(self, h, w):
        """``patch_or_image='orig'`` (reference inference.py:401-408): an engine whose workspace holds network inputs up
        to the image's own size rounded to multiples of 32 (grown in steps of 128 so that a dataset shares one engine)."""
        from .engine import closest_multiple_of
        need = max(closest_multiple_of(int(h)), closest_multiple_of(int(w)), 256)
        cap = (need + 127) // 128 * 128
        for key, e in self._engines.items():
            if isinstance(key, tuple) and key[0] == "orig" and key[1] >= need:
                return e
        e = OrderEngine(self.num_classes, cap, min(self.max_pairs, 64), self.device)
        e.load_state_dict(self._state)
        self._engines[("orig", cap)] = e
        return e

    def load_state_dict(self, sd):
        self._state = sd
        for e in self._engines.values():
            e.load_state_dict(sd)
        if self._trainer is not None:
            self._trainer.load_state_dict(sd)
        self._train_dirty = False

    def _sync_from_trainer(self):
        """After training steps the fp32 masters live in the TrainEngine: export them (reference layout) for the
        eval-mode engines and for save_state."""
        if self._trainer is not None and self._train_dirty:
            sd = self._trainer.state_dict()
            self._state = sd
            for e in self._engines.values():
                e.load_state_dict(sd)
            self._train_dirty = False

    def load_state(self, path, Iter=None, resume=False):
        """reference models/single_stage_model.py:54-61 + utils/common_utils.py:128-149."""
        if Iter is not None:
            path = os.path.join(path, "ckpt_iter_{}.pth.tar".format(Iter))
        if not os.path.isfile(path):
            raise Exception("=> no checkpoint found at '{}'".format(path))
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        self.load_state_dict(ckpt["state_dict"])
        if resume and ckpt.get("optimizer"):              # utils/common_utils.py:143-147
            self.optim.load_state_dict(ckpt["optimizer"])
        return ckpt["step"]

    def load_pretrain(self, load_path):
        self.load_state(load_path)

    def save_state(self, path, Iter):
        """reference models/single_stage_model.py:66-72 -- same file name and dict layout
        ({'step', 'state_dict' with the ``module.`` prefix, 'optimizer'})."""
        self._sync_from_trainer()
        if self._state is None:
            raise RuntimeError("no weights to save")
        path = os.path.join(path, "ckpt_iter_{}.pth.tar".format(Iter))
        sd = {(k if k.startswith("module.") else "module." + k): torch.as_tensor(np.asarray(v))
              for k, v in self._state.items()}
        torch.save({"step": Iter, "state_dict": sd, "optimizer": self.optim.state_dict()}, path)

    def switch_to(self, phase):
        """single_stage_model.py:74-78.  'eval' after training refreshes the eval-mode (folded-BN) engines."""
        if phase != "train":
            self._sync_from_trainer()
        self.phase = phase

    # ---- training step (reference models/supervised_order.py:83-95, 413-438, 481-493, 535-548) ----------------
    def _train_engine(self, batch, input_size):
        t = self._trainer
        if t is not None and (t.batch_pairs != batch or t.input_size != input_size):
            raise RuntimeError("the training engine was built for batches of [%d, %d^2]; got [%d, %d^2] "
                               "(the reference's DistributedGivenIterationSampler only yields full batches)" %
                               (t.batch_pairs, t.input_size, batch, input_size))
        if t is None:
            t = TrainEngine(self.num_classes, input_size, batch, self.device)
            t.load_state_dict(self._state)
            if self.dist_model:
                t.broadcast_params()                      # DistModule.__init__, utils/distributed_utils.py:17-24
            self.optim.attach(t)
            self._trainer = t
        return t

    def set_input_pairs(self, pair_tensor, input_size, *labels):
        """Batched alternative to ``set_input`` for the GPU data pipeline (instaorder_b200.train_data): the pair
        tensor of the batch as written by ``io_pair_gather_patch`` (both masks + normalised rgb, bf16) plus the same
        label tensors ``set_input`` takes.  The fp32 ``rgb / modal1 / modal2`` tensors of the reference never exist."""
        self._pair_tensor = (pair_tensor, int(input_size))
        dummy = torch.empty((self._n_from_labels(labels), 0, int(input_size), int(input_size)))
        self.set_input(dummy, dummy, dummy, *labels)

    @staticmethod
    def _n_from_labels(labels):
        return int(labels[0].shape[0])

    def _step(self, occ_off, class_off, class_k, occ_target, class_target, is_overlap):
        if self.phase != "train":
            raise RuntimeError("step() needs switch_to('train')")
        t = self._train_engine(int(self.rgb.shape[0]), int(self.rgb.shape[-1]))
        pt = getattr(self, "_pair_tensor", None)
        if pt is not None:
            self._pair_tensor = None
            if pt[0].data_ptr() != t.pair_tensor.data_ptr():
                t.pair_tensor[:pt[0].numel()].copy_(pt[0].view(torch.uint8).reshape(-1))
        else:
            t.pack_inputs(self.rgb, self.modal1, self.modal2)
        losses = t.forward_backward(occ_off, class_off, class_k, occ_target, class_target, is_overlap,
                                    float(self.params.get("overlap_weight", 1.0)),
                                    float(self.params.get("distinct_weight", 1.0)), self.world_size)
        out = losses.clone()
        self.optim.zero_grad()
        t.all_reduce_grads()                              # utils.average_gradients
        self.optim.step()
        self._train_dirty = True
        return out

    # ---- model(x): eval-mode forward of an arbitrary [B,5,D,D] batch (reference resnet_cls.py:203-222) ------
    @property
    def model(self):
        return _ModelCallable(self)

    def _forward_batch(self, rgb, modal1, modal2):
        """logits [B, 2, K] fp32 (both directions) for collated fp32 NCHW tensors on the device."""
        D = int(rgb.shape[-1])
        self._sync_from_trainer()
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
        raise NotImplementedError


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


def _check_labels(t, k, name):
    """Class targets must lie in [0, k): the reference's ``nn.CrossEntropyLoss`` raises on anything else (e.g. the -1
    the dataset classes emit for an unannotated depth pair); the loss kernels index with the target unchecked."""
    if t.numel():
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= k:
            raise IndexError("%s: target %d is out of bounds for %d classes" % (name, lo if lo < 0 else hi, k))
    return t


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

    def step(self):                                                               # supervised_order.py:535-548
        return {"loss": self._step(0, -1, 0, self.occ_order1, None, None)[0]}


class InstaOrderNet_d(_OrderModel):
    algo = "InstaOrderNet_d"

    def set_input(self, rgb=None, modal1=None, modal2=None, depth_order=None, count=None, is_overlap=None):  # :383-395
        self._set_common(rgb, modal1, modal2)
        self.depth_order1 = _check_labels(depth_order, 3, "depth_order").to(self.rgb.device, torch.int64).contiguous()
        self.depth_order2 = _swap01(self.depth_order1)
        self.count = count.to(self.rgb.device)
        self.is_overlap = is_overlap.to(self.rgb.device, torch.int64).contiguous()

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:397-411
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        if not ret_loss:
            return {}
        return {}, {"loss": self._loss(lg, -1, 0, 3, None, self.depth_order1, None)[0]}

    def step(self):                                           # supervised_order.py:413-438 (overlap / distinct weights)
        return {"loss": self._step(-1, 0, 3, None, self.depth_order1, self.is_overlap)[0]}


class InstaOrderNet_od(_OrderModel):
    algo = "InstaOrderNet_od"

    def set_input(self, rgb=None, modal1=None, modal2=None, depth_order=None, count=None, is_overlap=None,
                  occ_order=None):                                                # supervised_order.py:32-48
        self._set_common(rgb, modal1, modal2)
        self.depth_order1 = _check_labels(depth_order, 3, "depth_order").to(self.rgb.device, torch.int64).contiguous()
        self.depth_order2 = _swap01(self.depth_order1)
        self.count = count.to(self.rgb.device)
        self.is_overlap = is_overlap.to(self.rgb.device, torch.int64).contiguous()
        self.occ_order1 = occ_order.to(self.rgb.device, torch.float32).contiguous()
        self.occ_order2 = self.occ_order1[:, [1, 0]].contiguous()

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:50-81
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        out = self._loss(lg, 0, 2, 3, self.occ_order1, self.depth_order1, self.is_overlap)
        return {"loss_occ": out[1], "loss_depth": out[2]}, {"loss": out[0]}

    def step(self):                                                               # supervised_order.py:83-95
        out = self._step(0, 2, 3, self.occ_order1, self.depth_order1, self.is_overlap)
        return {"loss_occ": out[1], "loss_depth": out[2]}, {"loss": out[0]}


class OrderNet(_OrderModel):
    algo = "OrderNet"

    def set_input(self, rgb=None, modal1=None, modal2=None, occ_order=None):      # supervised_order.py:451-463
        self._set_common(rgb, modal1, modal2)
        self.occ_order1 = _check_labels(occ_order, int(self.num_classes), "occ_order").to(
            self.rgb.device, torch.int64).contiguous()
        self.occ_order2 = _swap01(self.occ_order1)

    def forward_only(self, ret_loss=True):                                        # supervised_order.py:465-479
        lg = self._forward_batch(self.rgb, self.modal1, self.modal2)
        if not ret_loss:
            return {}
        k = self.num_classes
        return {}, {"loss": self._loss(lg, -1, 0, int(k), None, self.occ_order1, None)[0]}

    def step(self):                                                               # supervised_order.py:481-493
        return {"loss": self._step(-1, 0, int(self.num_classes), None, self.occ_order1, None)[0]}


class InstaDepthNet_od(object):
    """Order inference with the reference's ``models.InstaDepthNet_od`` (models/supervised_order.py:99-237 wrapping
    midas/midas_net.py:113-212): ``load_state`` / ``load_state_dict`` / ``switch_to('eval')`` and the engine handle
    used by ``inference.infer_order_sup_occ_depth`` / ``infer_order_sup_depth`` (``method="InstaDepthNet_od"``).
    Inference only (order heads, disparity map, disparity-based depth order): training raises ``NotImplementedError``."""
    algo = "InstaDepthNet_od"
    with_occ = True
    with_trunks = True

    def __init__(self, params, load_pretrain=None, dist_model=False):
        self.params = params
        self.world_size = 1
        self.max_pairs = int(params.get("max_pairs", 64))
        self.max_images = int(params.get("max_images", 16))
        self.device = params.get("device", "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")))
        self._engines = {}
        self._state = None
        self.phase = "eval"
        if load_pretrain is not None:
            self.load_state(load_pretrain)

    def engine_for(self, input_size, disparity=False):
        """``disparity=True``: an engine whose encoder also runs layer4 and that can evaluate the MiDaS decoder."""
        from .depth_engine import DepthOrderEngine
        key = (input_size, bool(disparity))
        e = self._engines.get(key)
        if e is None:
            if self._state is None:
                raise RuntimeError("no weights loaded: call load_state()/load_state_dict() first")
            e = DepthOrderEngine(input_size, self.max_pairs, self.max_images, self.device, with_occ=self.with_occ,
                                 with_disparity=bool(disparity), with_trunks=self.with_trunks)
            e.load_state_dict(self._state)
            self._engines[key] = e
        return e

    def load_state_dict(self, sd):
        self._state = sd
        for e in self._engines.values():
            e.load_state_dict(sd)

    def load_state(self, path, Iter=None, resume=False):
        if Iter is not None:
            path = os.path.join(path, "ckpt_iter_{}.pth.tar".format(Iter))
        if not os.path.isfile(path):
            raise Exception("=> no checkpoint found at '{}'".format(path))
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        self.load_state_dict(ckpt["state_dict"])
        return ckpt.get("step", 0)

    def switch_to(self, phase):
        if phase == "train":
            raise NotImplementedError("InstaDepthNet: inference only")
        self.phase = phase

    def step(self):
        raise NotImplementedError("InstaDepthNet: inference only")


class InstaDepthNet_d(InstaDepthNet_od):
    """``models.InstaDepthNet_d`` (models/supervised_order.py:240-367 wrapping midas/midas_net.py:15-110): the same
    network without ``oo_net`` -- depth order only (``inference.infer_order_sup_depth``)."""
    algo = "InstaDepthNet_d"
    with_occ = False


class MidasNet(InstaDepthNet_od):
    """The plain MiDaS v2.1 network (reference midas/midas_net.py ``MidasNet``: ResNeXt-101 encoder + decoder, state_dict
    keys ``pretrained.*`` / ``scratch.*``) as used by ``infer_order_sup_depth(method="midas_pretrained")``
    (inference.py:576-583): disparity map and the median / mean depth order derived from it; no order heads."""
    algo = "midas_pretrained"
    with_occ = False
    with_trunks = False
