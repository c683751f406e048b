#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stem_pool.py tests/test_gpu_conv.py tests/test_gpu_order.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; tail -5 gpurun_out/r2c_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2c_layers.txt 2>&1; head -12 gpurun_out/r2c_layers.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:stem_pool_kernel|conv_halo_kernel" -s 2 -c 2 -o gpurun_out/r2c_new_kernels python tools/layer_report.py 256 > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
