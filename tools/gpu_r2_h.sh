#!/bin/bash
# CTA-pair conv kernel: parity tests, layer report with / without pairs, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py tests/test_gpu_train_kernels.py -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; tail -15 gpurun_out/r2h_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2h_layers_pair.txt 2>&1; head -42 gpurun_out/r2h_layers_pair.txt
INSTAORDER_PAIR=0 timeout 300 python tools/layer_report.py 256 > gpurun_out/r2h_layers_nopair.txt 2>&1; head -3 gpurun_out/r2h_layers_nopair.txt
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2h_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
timeout 600 python bench.py --workload train --steps 20 --no-cpu-baseline > gpurun_out/r2h_train.json 2> gpurun_out/r2h_train.err; python -c "
import json; j=json.load(open('gpurun_out/r2h_train.json')); print(j['value'], j['ms_per_step'], j['clocks'], j['roofline']['step_frac'])"
