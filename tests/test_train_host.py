"""Host-side logic of the training path on CPU: the flat-gradient all-reduce / parameter broadcast with a real
world_size-2 gloo group (checked against the training oracle's data-parallel semantics), the flat-buffer <-> reference
state_dict mapping rules, and the no-CPU-fallback guarantee."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instaorder_b200 import synth, training
from oracle import train_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALGO = "InstaOrderNet_o"


def _flat(grads, names):
    return torch.cat([torch.from_numpy(grads[k]).reshape(-1) for k in names])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nc = T.ALGOS[ALGO][0]
    names = T.param_names(nc)
    # rank 0 owns the "real" weights; the others start from garbage and receive the broadcast
    sd = synth.random_state_dict(3 if rank == 0 else 99, 5, nc)
    flat_w = torch.cat([torch.from_numpy(sd["module." + k]).reshape(-1) for k in names])
    training.broadcast_(flat_w, 0)
    off = 0
    for k in names:
        n = sd["module." + k].size
        sd["module." + k] = flat_w[off:off + n].reshape(sd["module." + k].shape).numpy().copy()
        off += n
    if rank != 0:      # running statistics are broadcast too (TrainEngine.broadcast_params)
        ref = synth.random_state_dict(3, 5, nc)
        for k in ref:
            if k.endswith(("running_mean", "running_var")):
                sd[k] = ref[k]
    batch = T.make_batch(50 + rank, 2, 64, ALGO)            # each rank its own slice of the global batch
    r = T.train_step(sd, batch, ALGO, world_size=world, apply_update=False)
    g = _flat(r["grads"], names)
    training.all_reduce_sum_(g)
    q.put((rank, float(r["loss"]), g.numpy(), float(flat_w.double().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_flat_allreduce_matches_data_parallel_semantics():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][3] == got[1][3]                            # broadcast: identical parameters on both ranks
    assert np.array_equal(got[0][2], got[1][2])              # all-reduce: identical gradients on both ranks
    # reference semantics (supervised_order.py:78 + distributed_utils.py:27-31): SUM over ranks of grad(loss_r / world)
    # = mean over ranks of the single-process gradients
    nc = T.ALGOS[ALGO][0]
    names = T.param_names(nc)
    sd = synth.random_state_dict(3, 5, nc)
    singles = []
    for rank in range(world):
        r = T.train_step(sd, T.make_batch(50 + rank, 2, 64, ALGO), ALGO, world_size=1, apply_update=False)
        singles.append(_flat(r["grads"], names).numpy())
        assert abs(r["loss"] / world - got[rank][1]) <= 1e-6 * abs(r["loss"])
    want = (singles[0] + singles[1]) / world
    err = float(np.abs(got[0][2] - want).max())
    assert err <= 1e-5 * float(np.abs(want).max()), err


def test_no_cpu_fallback_for_training():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        training.TrainEngine([2, 3], 64, 2)
    from instaorder_b200 import models
    m = models.InstaOrderNet_od(dict(algo="InstaOrderNet_od", backbone_arch="resnet50_cls",
                                     backbone_param=dict(in_channels=5, num_classes=[2, 3]), optim="SGD", lr=1e-4,
                                     weight_decay=1e-4))
    assert m.optim.param_groups[0]["lr"] == 1e-4 and m.optim.param_groups[0]["momentum"] == 0.9
    m.optim.param_groups[0]["lr"] = 5e-5                     # what utils.StepLRScheduler does (scheduler.py:77-80)
    assert m.optim.param_groups[0]["lr"] == 5e-5
    with pytest.raises(Exception):
        models.InstaOrderNet_od(dict(algo="InstaOrderNet_od", optim="LAMB", backbone_param=dict(num_classes=[2, 3])))


def test_param_order_matches_reference_parameters():
    """FlatOptim.state_dict indexes parameters like torch.optim does: model.parameters() order."""
    names = T.param_names([2, 3])
    assert names[0] == "conv1.weight" and names[1] == "bn1.weight" and names[2] == "bn1.bias"
    assert names[-4:] == ["fc_occ.weight", "fc_occ.bias", "fc_depth.weight", "fc_depth.bias"]
    assert len(names) == 163                                       # 53 convs + 106 BN affine + 4 FC tensors
    n = sum(int(np.prod(s)) for k, s in synth.resnet50_layout(5, [2, 3]) if k in names)
    assert n == 23524549                                           # 23,524,549 parameters
