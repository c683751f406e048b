#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order.py tests/test_gpu_metrics.py tests/test_gpu_gather.py tests/test_gpu_instadepth.py -m gpu -x -q -s > gpurun_out/r2d_tests.log 2>&1; grep -E "orig head|passed|failed|Error" gpurun_out/r2d_tests.log | tail -12
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2d_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks']); print(j['roofline']['metrics'])"
