#!/bin/bash
# Round-2 final pass C (2 GPUs): the bench exactly as the driver launches it, the reference arm under torchrun, DDP check.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2w_bench_2gpu.json 2> gpurun_out/r2w_bench_2gpu.err; python -c "
import json; j=json.load(open('gpurun_out/r2w_bench_2gpu.json')); print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['training']['value'], j['training']['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2w_ref_2gpu.json 2> gpurun_out/r2w_ref_2gpu.err; cut -c1-200 gpurun_out/r2w_ref_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/train_ddp_check.py > gpurun_out/r2w_ddp_check.txt 2>&1; tail -4 gpurun_out/r2w_ddp_check.txt
