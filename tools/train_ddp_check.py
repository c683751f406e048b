"""2-rank data-parallel check of the training step on real GPUs (run under torch.distributed.run, NCCL):
  * parameters are broadcast from rank 0 (the ranks start from different seeds on purpose),
  * every rank steps on ITS OWN batch; after the flat-gradient all-reduce + SGD the parameters are bit-identical on all
    ranks (same summed gradient, same update),
  * the all-reduced gradient equals the sum of the per-rank gradients (recomputed on rank 0 from both batches)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instaorder_b200 import models, synth  # noqa: E402
from oracle import gen_golden_train as G  # noqa: E402
from oracle import train_oracle as T  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    algo, B, D = "InstaOrderNet_od", 4, 128
    params = dict(algo=algo, backbone_arch="resnet50_cls", backbone_param=dict(in_channels=5, num_classes=[2, 3]),
                  optim="SGD", lr=1e-2, weight_decay=1e-4, use_rgb=True, overlap_weight=0.1, distinct_weight=0.9,
                  device=dev)
    m = models.InstaOrderNet_od(params, dist_model=True)
    m.load_state_dict(synth.random_state_dict(40 + rank, 5, [2, 3]))     # different weights per rank before broadcast
    m.switch_to("train")
    ok = True
    for it in range(3):
        batch = T.make_batch(1000 * rank + it, B, D, algo)
        m.set_input(**G.set_input_args(algo, batch))
        log, out = m.step()
        eng = m._trainer
        cs = torch.stack([eng.params.double().sum(), eng.params.double().abs().sum(), eng.grads.double().sum(),
                          eng.stats.double().sum()])
        all_cs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(all_cs, cs)
        same_params = all(torch.equal(all_cs[0][:2], c[:2]) for c in all_cs)
        same_grads = all(torch.equal(all_cs[0][2], c[2]) for c in all_cs)
        if rank == 0:
            print("step %d: loss %.5f  params identical across ranks: %s  reduced grads identical: %s  (running stats "
                  "are per-rank: %s)" % (it, float(out["loss"]), same_params, same_grads,
                                         [float(c[3]) for c in all_cs]))
        ok = ok and same_params and same_grads
    # bucketed / overlapped all-reduce against the plain one: the same backward pass twice (no optimiser step in
    # between) -- (a) synchronise, then ONE all-reduce of the whole buffer; (b) the four bucket all-reduces issued while
    # backward is still running.  The two gradients agree up to the run-to-run order of the fp32 atomics in the
    # weight-gradient kernels; a bucket reduced before it was complete would be off by O(1).
    eng = m._trainer
    batch = T.make_batch(5000 + rank, B, D, algo)
    m.set_input(**G.set_input_args(algo, batch))
    eng.pack_inputs(m.rgb, m.modal1, m.modal2)
    args = (0, 2, 3, m.occ_order1, m.depth_order1, m.is_overlap, 0.1, 0.9, world)
    eng.forward_backward(*args)
    torch.cuda.synchronize()
    want = eng.grads.clone()
    dist.all_reduce(want)
    eng.forward_backward(*args)
    eng.all_reduce_grads()
    torch.cuda.synchronize()
    worst = 0.0
    for (b0, e0) in eng.buckets():
        scale = float(want[b0:e0].abs().max()) + 1e-20
        worst = max(worst, float((eng.grads[b0:e0] - want[b0:e0]).abs().max()) / scale)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bucket_ok = float(t) < 1e-3
    ok = ok and bucket_ok
    if rank == 0:
        print("bucketed all-reduce vs single all-reduce of a synchronised backward: max |diff| / bucket max = %.3g (%s); "
              "buckets %s" % (float(t), "ok" if bucket_ok else "MISMATCH", eng.buckets()))
    if rank == 0:
        eng = m._trainer
        # two runs of this script with INSTAORDER_ALLREDUCE_OVERLAP=1 / 0 must print the same line (2 ranks: the sum of
        # two gradients does not depend on how the buffer is cut into all-reduce calls)
        print("FINAL overlap=%s params_sum=%.17g params_abs=%.17g grads_abs=%.17g" % (
            os.environ.get("INSTAORDER_ALLREDUCE_OVERLAP", "1"), float(eng.params.double().sum()),
            float(eng.params.double().abs().sum()), float(eng.grads.double().abs().sum())))
        print("DDP CHECK", "PASSED" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
