"""Training / validation driver with the reference's ``Trainer`` interface (reference trainer.py:24-266): same config
sections (``args.model`` / ``args.data`` / ``args.trainer``), same iteration-based loop -- LR schedule before every
step, ``set_input(*inputs)`` + ``step()``, loss averaged over ranks for logging, ``save_state`` / ``validate`` cadence
-- driving the B200 models of ``instaorder_b200.models``.

Differences that are deliberate:
  * the dataset objects are injected (``train_dataset`` / ``val_dataset``): the reference's COCO / KINS readers
    (datasets/*.py, pycocotools) are outside this path; anything whose ``__getitem__`` returns the reference's sample
    tuple works, e.g. ``SyntheticPairDataset`` below;
  * logging goes to the python logger / a log file (wandb and tensorboardX are optional and skipped when absent);
  * samplers and the LR schedule are the pure functions of ``instaorder_b200.sharding`` (pinned to the reference's
    classes by tests/test_sharding.py).
"""
import logging
import os
import time
from datetime import datetime

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Sampler

from . import models, sharding


class AverageMeter(object):
    """reference utils/common_utils.py:81-109."""

    def __init__(self, length=0):
        self.length = length
        self.reset()

    def reset(self):
        if self.length > 0:
            self.history = []
        else:
            self.count = 0
            self.sum = 0.0
        self.val = 0.0
        self.avg = 0.0

    def update(self, val):
        if self.length > 0:
            self.history.append(val)
            if len(self.history) > self.length:
                del self.history[0]
            self.val = self.history[-1]
            self.avg = np.mean(self.history)
        else:
            self.val = val
            self.sum += val
            self.count += 1
            self.avg = self.sum / self.count


class _ListSampler(Sampler):
    def __init__(self, indices):
        self.indices = list(int(i) for i in indices)

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


class StepLRScheduler(object):
    """utils/scheduler.py:84-109 over ``optimizer.param_groups`` (our FlatOptim or a torch optimiser)."""

    def __init__(self, optimizer, milestones, lr_mults, base_lr, warmup_lr, warmup_steps, last_iter=-1):
        assert len(milestones) == len(lr_mults), "{} vs {}".format(milestones, lr_mults)
        self.optimizer = optimizer
        self.kw = dict(base_lr=base_lr, milestones=list(milestones), lr_mults=list(lr_mults),
                       warmup_lr=list(warmup_lr), warmup_steps=list(warmup_steps))
        self.last_iter = last_iter

    def step(self, this_iter=None):
        self.last_iter = self.last_iter + 1 if this_iter is None else this_iter
        lr = sharding.step_lr(self.last_iter, **self.kw)
        for g in self.optimizer.param_groups:
            g["lr"] = lr

    def get_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]


def reduce_tensors(t):
    """utils/distributed_utils.py:133-136: SUM over ranks (the losses are already divided by world_size)."""
    import torch.distributed as dist
    r = t.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(r)
    return r


class SyntheticPairDataset(Dataset):
    """Seeded synthetic samples in the reference ``__getitem__`` formats:
    ^od: (rgb[3,D,D], modal1[1,D,D], modal2[1,D,D], depth_order, count, is_overlap, occ_order[2]);
    ^d: without occ_order; ^o: (rgb, modal1, modal2, occ_order[2] float); OrderNet: (..., occ_order int64)."""

    def __init__(self, algo, input_size, length, seed=0):
        self.algo, self.D, self.length, self.seed = algo, int(input_size), int(length), int(seed)

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        rng = np.random.RandomState(self.seed * 1000003 + int(idx))
        D = self.D
        rgb = torch.from_numpy(rng.standard_normal((3, D, D)).astype(np.float32))
        yy, xx = np.mgrid[0:D, 0:D]
        ms = []
        for _ in range(2):
            cx, cy = rng.uniform(0.2, 0.8, 2) * D
            rx, ry = rng.uniform(0.1, 0.4, 2) * D
            ms.append(torch.from_numpy((((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0).astype(np.float32))[None])
        depth = torch.tensor(int(rng.randint(0, 3)), dtype=torch.int64)
        count = torch.tensor(int(rng.randint(2, 4)), dtype=torch.int64)
        ovl = torch.tensor(int(rng.rand() < 0.3), dtype=torch.int64)
        occ2 = torch.from_numpy((rng.rand(2) < 0.3).astype(np.float32))
        if self.algo == "InstaOrderNet_od":
            return rgb, ms[0], ms[1], depth, count, ovl, occ2
        if self.algo == "InstaOrderNet_d":
            return rgb, ms[0], ms[1], depth, count, ovl
        if self.algo == "InstaOrderNet_o":
            return rgb, ms[0], ms[1], occ2
        return rgb, ms[0], ms[1], torch.tensor(int(rng.randint(0, 3)), dtype=torch.int64)


class Trainer(object):
    def __init__(self, args, train_dataset=None, val_dataset=None):
        import torch.distributed as dist
        seed = getattr(args, "seed", 0)
        torch.manual_seed(seed)            # reference trainer.py:26-31 (the model constructor below draws its
        np.random.seed(seed)               # initial weights from torch's generator right after this seeding)
        import random
        random.seed(seed)
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        self.args = args
        self.logger = None
        if self.rank == 0:
            now = datetime.now().strftime("%m-%d-%H-%M")
            self.run_name = "%s_%s_%s_%s" % (args.data.get("dataset", "synthetic"), args.trainer.get("exp_name", "exp"),
                                             args.data.get("patch_or_image", "patch"), now)
            base = args.data.get("base_dir", ".")
            self.folder2save = os.path.join(base, "data", "out", "InstaOrder", self.run_name)
            os.makedirs(os.path.join(self.folder2save, "logs"), exist_ok=True)
            os.makedirs(os.path.join(self.folder2save, "checkpoints"), exist_ok=True)
            self.logger = logging.getLogger("global_logger_%s" % self.run_name)
            self.logger.setLevel(logging.INFO)
            fh = logging.FileHandler(os.path.join(self.folder2save, "logs",
                                                  "log_offline_val.txt" if getattr(args, "validate", False)
                                                  else "log_train.txt"))
            fh.setFormatter(logging.Formatter("[%(asctime)s] %(message)s"))
            self.logger.addHandler(fh)
        # model (reference trainer.py:84)
        self.model = models.__dict__[args.model["algo"]](args.model, load_pretrain=getattr(args, "load_pretrain", None),
                                                         dist_model=self.distributed)
        self.start_iter = 0
        if getattr(args, "load_model", None) is not None:
            self.model.load_state(args.load_model, Iter=None, resume=True)
            self.start_iter = int(args.load_model.split("iter_")[-1].split(".")[0])
        self.curr_step = self.start_iter
        bs = args.data["batch_size"]
        if not getattr(args, "validate", False):
            self.lr_scheduler = StepLRScheduler(self.model.optim, args.model["lr_steps"], args.model["lr_mults"],
                                                args.model["lr"], args.model["warmup_lr"], args.model["warmup_steps"],
                                                last_iter=self.start_iter - 1)
            if train_dataset is None:
                raise ValueError("pass a train_dataset (the reference's COCO / KINS readers are out of this path's scope)")
            idx = sharding.given_iteration_indices(len(train_dataset), args.model["total_iter"], bs, self.rank,
                                                   self.world_size, last_iter=self.start_iter - 1)
            self.train_loader = DataLoader(train_dataset, batch_size=bs, shuffle=False,
                                           num_workers=args.data.get("workers", 0), pin_memory=True,
                                           sampler=_ListSampler(idx))
        self.val_loader = None
        if val_dataset is not None:
            vidx = sharding.sequential_indices(len(val_dataset), self.rank, self.world_size)
            self.val_loader = DataLoader(val_dataset, batch_size=args.data.get("batch_size_val", bs), shuffle=False,
                                         num_workers=args.data.get("workers", 0), pin_memory=True,
                                         sampler=_ListSampler(vidx))
        self.history = []          # (step, {name: averaged loss}) -- what the reference sends to wandb

    def run(self):
        if getattr(self.args, "validate", False):
            self.validate("off_val")
            return
        if self.args.trainer.get("initial_val", False):
            self.validate("on_val")
        self.train()

    def _log(self, msg):
        if self.rank == 0 and self.logger is not None:
            self.logger.info(msg)

    def train(self):
        args = self.args
        btime_rec, dtime_rec = AverageMeter(10), AverageMeter(10)
        recorder = {rec: AverageMeter(10) for rec in args.trainer["loss_record"]}
        self.model.switch_to("train")
        end = time.time()
        total_iter = args.model["total_iter"]
        for i, inputs in enumerate(self.train_loader):
            self.curr_step = self.start_iter + i
            self.lr_scheduler.step(self.curr_step)
            curr_lr = self.lr_scheduler.get_lr()[0]
            dtime_rec.update(time.time() - end)
            self.model.set_input(*inputs)
            loss_dict = self.model.step()
            loss_dict_to_log = {}
            if len(loss_dict) == 2:
                loss_dict_to_log, loss_dict = loss_dict
            for k in loss_dict.keys():
                recorder[k].update(reduce_tensors(loss_dict[k]).item())
            btime_rec.update(time.time() - end)
            end = time.time()
            self.curr_step += 1
            if self.rank == 0 and self.curr_step % args.trainer["print_freq"] == 0:
                self.history.append((self.curr_step, {k: float(r.avg) for k, r in recorder.items()}))
                loss_str = "".join("{}: {:.4g} ({:.4g})\t".format(k, r.val, r.avg) for k, r in recorder.items())
                self._log("Iter: [{0}/{1}]\tTime {2:.3f} ({3:.3f})\tData {4:.3f} ({5:.3f})\t{6}lr {7:.2g}".format(
                    self.curr_step, len(self.train_loader), btime_rec.val, btime_rec.avg, dtime_rec.val,
                    dtime_rec.avg, loss_str, curr_lr))
            if self.rank == 0 and (self.curr_step % args.trainer["save_freq"] == 0 or self.curr_step == total_iter):
                self.model.save_state(os.path.join(self.folder2save, "checkpoints"), self.curr_step)
            if self.val_loader is not None and (self.curr_step % args.trainer["val_freq"] == 0 or
                                                self.curr_step == total_iter):
                self.validate("on_val")

    def validate(self, phase):
        args = self.args
        recorder = {rec: AverageMeter(10) for rec in args.trainer["loss_record"]}
        if self.val_loader is None:
            raise ValueError("validation requested (args.validate / initial_val / val_freq) but no val_dataset was "
                             "passed to Trainer(args, train_dataset, val_dataset)")
        self.model.switch_to("eval")
        for i, inputs in enumerate(self.val_loader):
            if args.trainer.get("val_iter", -1) != -1 and i == args.trainer["val_iter"]:
                break
            self.model.set_input(*inputs)
            loss_dict_to_log, loss_dict = self.model.forward_only()
            for k in loss_dict.keys():
                recorder[k].update(reduce_tensors(loss_dict[k]).item())
        if self.rank == 0:
            self.history.append((self.curr_step, {"val_" + k: float(r.avg) for k, r in recorder.items()}))
            self._log("Validation Iter: [{0}]\t".format(self.curr_step) +
                      "".join("{}: {:.4g} ({:.4g})\t".format(k, r.val, r.avg) for k, r in recorder.items()))
        self.model.switch_to("train")
