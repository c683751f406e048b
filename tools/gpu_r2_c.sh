#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem_pool.py tests/test_gpu_order.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; tail -5 gpurun_out/r2c_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2c_layers.txt 2>&1; head -5 gpurun_out/r2c_layers.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_pool_kernel -s 2 -c 1 -o gpurun_out/r2c_stem_pool python tools/layer_report.py 256 > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
