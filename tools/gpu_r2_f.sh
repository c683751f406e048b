#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ov in 1 0; do
  INSTAORDER_ALLREDUCE_OVERLAP=$ov timeout 300 $TR tools/train_ddp_check.py > gpurun_out/r2f_ddp_overlap$ov.txt 2>&1; grep -E "bucketed|FINAL|DDP CHECK|Error" gpurun_out/r2f_ddp_overlap$ov.txt
done
