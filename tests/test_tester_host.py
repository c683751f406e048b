"""Multi-rank host logic of the evaluation driver (instaorder_b200/tester.py) on CPU: a real world_size-2 gloo group,
images sharded round-robin, per-image rows gathered, the dataset-level numbers identical on both ranks and equal to a
single-process evaluation.  The engine is a stand-in that returns deterministic matrices, and the two metric entry
points are served by the CPU oracle (the CUDA kernels are checked by the -m gpu tests)."""
import os
import sys
import types

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEngine(object):
    device = "cpu"

    def infer_scenes(self, scenes, method, pairs="all", patch_or_image="patch"):
        out = []
        for sc in scenes:
            rng = np.random.RandomState(int(sc.image[0, 0, 0]) + 100 * sc.n)
            occ = (rng.rand(sc.n, sc.n) < 0.3).astype(np.int64)
            depth = rng.randint(0, 3, size=(sc.n, sc.n)).astype(np.int64)
            out.append(dict(occ=occ, depth=depth))
        return out


class _Reader(object):
    def __init__(self, n):
        from instaorder_b200 import synth
        rng = np.random.RandomState(1)
        self.items = []
        for k in range(n):
            m = int(rng.randint(2, 6))
            image = np.full((8, 8, 3), k, np.uint8)
            masks = np.zeros((m, 8, 8), np.uint8)
            occ, depth, overlap, count = synth.make_gt(rng, m)
            self.items.append((image, masks, np.tile(np.array([[0., 0., 4., 4.]]), (m, 1)), occ, depth, overlap, count))

    def __len__(self):
        return len(self.items)

    def get_image_instances(self, i, with_gt=False):
        it = self.items[i]
        return it[1], np.ones(len(it[1]), np.int64), it[2], np.array([]), "img_%d" % i

    def get_gt_ordering(self, i, kind, rm_bidirec=0):
        it = self.items[i]
        return [it[4], it[5], it[6]] if kind == "depth" else it[3]


def _evaluate(n_images):
    from instaorder_b200 import engine, tester
    from oracle import oracle as O
    # metric kernels -> CPU oracle (same float64 definitions, bit-identical by tests/test_gpu_metrics.py)
    engine.metrics_prf = lambda orders, gts, zd, device=None: np.array(
        [O.eval_order_recall_precision_f1(o, g, zd) for o, g in zip(orders, gts)], np.float64)
    engine.metrics_whdr = lambda orders, gt, ov, ct, device=None: np.array(
        [[O.eval_depth_order_whdr(o, (a, b, c))[k][0] for k in tester._infer.WHDR_KEYS] for o, a, b, c in zip(orders, gt, ov, ct)],
        np.float64)
    engine.expand_bbox = lambda b, e=3.0: np.asarray(b)
    reader = _Reader(n_images)
    model = types.SimpleNamespace(engine_for=lambda d: _FakeEngine())
    args = types.SimpleNamespace(order_method="InstaOrderNet_od", pairs="all", zd=1, disp_select_method="",
                                 data=dict(patch_or_image="patch", input_size=256, remove_occ_bidirec=0), images_per_call=2)
    return tester.Tester(args, model, reader, lambda fn: reader.items[int(fn.split("_")[1])][0]).run()


def _worker(rank, world, port, n_images, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = _evaluate(n_images)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_tester_two_ranks_equal_single_process():
    n_images = 9
    single = _evaluate(n_images)
    assert single["val/num_test_images"] == n_images and "val_ovlX/WHDR_all" in single and "val/f1" in single
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r].keys() == single.keys()
        for k, v in single.items():
            assert got[r][k] == v, (r, k, got[r][k], v)
