#!/bin/bash
# conv_row3 (stacked-kx 3x3), residual via MMA + relu-pack in conv_fused: parity tests, layer report, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; tail -15 gpurun_out/r2j_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2j_layers.txt 2>&1; head -42 gpurun_out/r2j_layers.txt
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2j_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
INSTAORDER_RES_MMA=0 INSTAORDER_ROW3=0 timeout 300 python tools/layer_report.py 256 > gpurun_out/r2j_layers_base.txt 2>&1; head -12 gpurun_out/r2j_layers_base.txt
