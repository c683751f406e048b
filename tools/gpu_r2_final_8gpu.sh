#!/bin/bash
# Round-2 final pass D (8 GPUs): the bench as the driver launches it (inference line + nested training line).
mkdir -p gpurun_out
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2v_bench_8gpu.json 2> gpurun_out/r2v_bench_8gpu.err; python -c "
import json; j=json.load(open('gpurun_out/r2v_bench_8gpu.json')); print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['per_call']['value'], j['clocks'], j['training']['value'], j['training']['ms_per_step'])"
tail -3 gpurun_out/r2v_bench_8gpu.err
