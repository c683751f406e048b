#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stem_pool.py -m gpu -x -q > gpurun_out/r2b_stem.log 2>&1; tail -15 gpurun_out/r2b_stem.log
timeout 600 python -m pytest tests/test_gpu_order.py -m gpu -x -q > gpurun_out/r2b_order.log 2>&1; tail -5 gpurun_out/r2b_order.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2b_layers.txt 2>&1; head -8 gpurun_out/r2b_layers.txt
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
