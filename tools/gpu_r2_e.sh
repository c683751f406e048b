#!/bin/bash
# 2-GPU: bucketed / overlapped gradient all-reduce vs the single call (correctness + timing), orig mode, whdr kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_order.py tests/test_gpu_metrics.py tests/test_gpu_instadepth.py tests/test_gpu_train_step.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; tail -3 gpurun_out/r2e_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ov in 1 0; do
  INSTAORDER_ALLREDUCE_OVERLAP=$ov timeout 300 $TR tools/train_ddp_check.py > gpurun_out/r2e_ddp_overlap$ov.txt 2>&1; grep -E "FINAL|DDP CHECK" gpurun_out/r2e_ddp_overlap$ov.txt
done
for ov in 1 0; do
  INSTAORDER_ALLREDUCE_OVERLAP=$ov timeout 300 $TR bench.py --gpus 2 --workload train --steps 30 --warmup 5 > gpurun_out/r2e_train_2gpu_overlap$ov.json 2> gpurun_out/r2e_train_2gpu_overlap$ov.err
  python -c "
import json; j=json.load(open('gpurun_out/r2e_train_2gpu_overlap$ov.json')); print('overlap=$ov', j['value'], j['ms_per_step'], j['clocks'])"
done
timeout 300 python bench.py --workload train --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_train_1gpu.json 2> gpurun_out/r2e_train_1gpu.err; python -c "
import json; j=json.load(open('gpurun_out/r2e_train_1gpu.json')); print('1gpu', j['value'], j['ms_per_step'], j['clocks'])"
