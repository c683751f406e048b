#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; tail -3 gpurun_out/r2l_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2l_layers.txt 2>&1; head -42 gpurun_out/r2l_layers.txt
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2l_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
