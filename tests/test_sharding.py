"""Host-side multi-rank logic on CPU: image sharding + the one metrics gather, with a real world_size-2 gloo group,
and the reference's sampler / LR-schedule index rules."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instaorder_b200 import sharding
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_images, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_interleaved(n_images, rank, world)
    # every rank computes "metrics" only for its own images: row = f(image index)
    rows = [[i * 1.5, i + 0.25, -1.0 if i % 3 == 0 else i * 2.0] for i in mine]
    table = sharding.gather_metric_rows(rows, mine, n_images)
    q.put((rank, mine, table))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_metric_gather():
    world, n_images = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(sum((g[1] for g in got), []))
    assert owned == list(range(n_images))                       # a partition: every image exactly once
    want = np.array([[i * 1.5, i + 0.25, -1.0 if i % 3 == 0 else i * 2.0] for i in range(n_images)])
    for _, _, table in got:
        assert np.array_equal(table, want)                      # every rank ends with the full table, in order


def test_aggregate_metrics_matches_reference_formula():
    rng = np.random.RandomState(0)
    prf = rng.rand(11, 3) * 100
    whdr = rng.rand(11, 9) * 100
    whdr[rng.rand(11, 9) < 0.3] = -1
    out = sharding.aggregate_metrics(prf, whdr)
    assert out["recall"] == sum(prf[:, 0].tolist()) / 11            # tools/test.py:274-276
    for k, key in enumerate(sharding.WHDR_KEYS):
        col = whdr[:, k]
        v = col[col != -1]
        assert out["WHDR_" + key] == v.sum() / (len(v) + 1e-6)      # tools/test.py:266-268


def test_sampler_index_rules():
    assert sharding.sequential_indices(10, 0, 4) == [0, 1, 2]
    assert sharding.sequential_indices(10, 3, 4) == [9, 0, 1]       # wrap-padded tail block
    a = sharding.given_iteration_indices(50, total_iter=7, batch_size=4, rank=0, world=2)
    b = sharding.given_iteration_indices(50, total_iter=7, batch_size=4, rank=1, world=2)
    assert len(a) == len(b) == 28 and not np.array_equal(a, b)
    r = sharding.given_iteration_indices(50, 7, 4, 0, 2, last_iter=2)
    assert np.array_equal(r, a[12:])                                # resume skips (last_iter + 1) * batch


def test_step_lr_schedule():
    # experiments/InstaOrder/InstaOrderNet_od/config.yaml: lr 1e-4, steps [32000, 48000], mults [0.1, 0.1]
    kw = dict(base_lr=1e-4, milestones=[32000, 48000], lr_mults=[0.1, 0.1], warmup_lr=[], warmup_steps=[])
    assert sharding.step_lr(0, **kw) == 1e-4
    assert sharding.step_lr(31999, **kw) == 1e-4
    assert abs(sharding.step_lr(32000, **kw) - 1e-5) < 1e-18
    assert abs(sharding.step_lr(50000, **kw) - 1e-6) < 1e-18
    kw = dict(base_lr=0.1, milestones=[100], lr_mults=[0.5], warmup_lr=[0.4], warmup_steps=[10])
    assert abs(sharding.step_lr(5, **kw) - 0.25) < 1e-12
    assert abs(sharding.step_lr(10, **kw) - 0.4) < 1e-12
    assert abs(sharding.step_lr(100, **kw) - 0.2) < 1e-12


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_samplers_and_scheduler_against_the_reference():
    ns = ref_shim.load()
    du = ns.utils
    ds = list(range(37))
    for world in (1, 2, 4):
        for rank in range(world):
            s = du.DistributedSequentialSampler(ds, world, rank)
            assert list(iter(s)) == sharding.sequential_indices(len(ds), rank, world)
            g = du.DistributedGivenIterationSampler(ds, 5, 3, world, rank, last_iter=1)
            assert list(iter(g)) == list(sharding.given_iteration_indices(len(ds), 5, 3, rank, world, last_iter=1))
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
    sch = du.StepLRScheduler(opt, [20, 30], [0.1, 0.5], 0.1, [0.2, 0.4], [5, 10], last_iter=-1)
    for it in range(40):
        sch.step(it)
        want = sharding.step_lr(it, 0.1, [20, 30], [0.1, 0.5], [0.2, 0.4], [5, 10])
        assert abs(opt.param_groups[0]["lr"] - want) < 1e-15, it
