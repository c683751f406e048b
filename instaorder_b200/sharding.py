"""Multi-GPU plumbing of the pairwise-order path: images are the unit of sharding (every image's pairs, order
matrices and metrics are independent), so ranks never exchange activations; only the per-image metric rows are
merged at the end (reference tools/test.py:264-283, M3 in SURVEY.md section 8a).

Also the index lists of the reference's distributed samplers (utils/distributed_utils.py:139-160, 203-254) and its
iteration-based LR schedule (utils/scheduler.py:58-109), as pure functions, for the training driver.
"""
import math
from bisect import bisect_right

import numpy as np

WHDR_KEYS = ["%s_%s" % (o, e) for o in ("ovlX", "ovlO", "ovlOX") for e in ("eq", "neq", "all")]


def shard_interleaved(n_items, rank, world):
    """Round-robin image sharding used by the inference driver / bench: rank r owns r, r + world, ..."""
    return list(range(rank, n_items, world))


def sequential_indices(n_items, rank, world):
    """DistributedSequentialSampler (reference utils/distributed_utils.py:139-160): contiguous block of
    ceil(n / world) indices per rank, wrap-padded from the start of the dataset."""
    assert n_items >= world, "{} vs {}".format(n_items, world)
    sub = int(math.ceil(n_items * 1.0 / world))
    padded = list(range(n_items)) + list(range(sub * world - n_items))
    return [padded[i] for i in range(sub * rank, sub * rank + sub)]


def given_iteration_indices(n_items, total_iter, batch_size, rank, world, last_iter=-1):
    """DistributedGivenIterationSampler (reference utils/distributed_utils.py:203-254): every rank shuffles the same
    tiled index list with ``np.random.seed(0)`` and takes its slice; a resumed run skips (last_iter+1)*batch."""
    total = total_iter * batch_size
    all_size = total * world
    idx = np.arange(n_items)[:all_size]
    idx = np.tile(idx, (all_size - 1) // idx.shape[0] + 1)[:all_size]
    st = np.random.get_state()
    np.random.seed(0)
    np.random.shuffle(idx)
    np.random.set_state(st)
    idx = idx[total * rank: total * rank + total]
    return idx[(last_iter + 1) * batch_size:]


def step_lr(it, base_lr, milestones, lr_mults, warmup_lr=(), warmup_steps=()):
    """StepLRScheduler (reference utils/scheduler.py:58-109) evaluated at global iteration ``it``."""
    warmup_lr, warmup_steps = list(warmup_lr), list(warmup_steps)
    pos = bisect_right(warmup_steps, it)
    if pos < len(warmup_steps):
        if pos == 0:
            cur = base_lr + it * (warmup_lr[pos] - base_lr) / warmup_steps[pos]
        else:
            cur = warmup_lr[pos - 1] + (it - warmup_steps[pos - 1]) * (warmup_lr[pos] - warmup_lr[pos - 1]) / (
                warmup_steps[pos] - warmup_steps[pos - 1])
        return base_lr * (cur / base_lr)
    mults = [1.0]
    for x in lr_mults:
        mults.append(mults[-1] * x)
    pos = bisect_right(list(milestones), it)
    scale = mults[pos] if len(warmup_lr) == 0 else warmup_lr[-1] * mults[pos] / base_lr
    return base_lr * scale


def aggregate_metrics(prf_rows, whdr_rows):
    """Dataset-level numbers exactly as tools/test.py:264-283 computes them: P/R/F1 = plain mean over images;
    WHDR per key = sum / (count + 1e-6) over the images whose value is not -1."""
    out = {}
    if prf_rows is not None and len(prf_rows):
        a = np.asarray(prf_rows, dtype=np.float64)
        out["recall"] = sum(a[:, 0].tolist()) / len(a)
        out["precision"] = sum(a[:, 1].tolist()) / len(a)
        out["f1"] = sum(a[:, 2].tolist()) / len(a)
        out["num_test_images"] = len(a)
    if whdr_rows is not None and len(whdr_rows):
        w = np.asarray(whdr_rows, dtype=np.float64)
        for k, key in enumerate(WHDR_KEYS):
            col = w[:, k]
            valid = col != -1
            out["WHDR_" + key] = col[valid].sum() / (len(col[valid]) + 1e-6)
    return out


def gather_metric_rows(rows, indices, n_total, group=None):
    """All ranks contribute float64 rows for the images they own (``indices``); returns the [n_total, C] table in
    dataset order on every rank.  One small all_gather of padded tensors -- the only collective of the inference
    path.  Works with gloo (CPU) and nccl."""
    import torch
    import torch.distributed as dist
    rows = np.asarray(rows, dtype=np.float64).reshape(len(indices), -1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = np.full((n_total, rows.shape[1]), np.nan)
        out[np.asarray(indices, dtype=np.int64)] = rows
        return out
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    cap = int(math.ceil(n_total / world)) + 1
    c = rows.shape[1]
    buf = torch.full((cap, c + 1), -1.0, dtype=torch.float64, device=dev)
    if len(indices):
        buf[: len(indices), 0] = torch.as_tensor(np.asarray(indices, dtype=np.float64), device=dev)
        buf[: len(indices), 1:] = torch.as_tensor(rows, device=dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.full((n_total, c), np.nan)
    for p in parts:
        p = p.cpu().numpy()
        keep = p[:, 0] >= 0
        out[p[keep, 0].astype(np.int64)] = p[keep, 1:]
    return out
